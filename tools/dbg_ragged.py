"""Developer: gradient errors of the smpl pipeline at ragged vs aligned sample counts (tests/test_gpu_train.py::_grad_check)."""
import sys
sys.path.insert(0, '/root/repo')
from tests import test_gpu_train as T
for shape in ((5, 7, 24, 37), (4, 8, 24, 40), (5, 7, 32, 29), (8, 8, 24, 37)):
    for kind in ('smpl',):
        try:
            w = T._grad_check(kind, 0, 1e-3, shape=shape)
            print(shape, kind, 'ok worst', w)
        except AssertionError as e:
            print(shape, kind, 'FAIL', str(e)[:200])
