/* Plain-C restatement of torchsearchsorted's row-wise bisection -- TEST INFRASTRUCTURE ONLY.
 *
 * Follows torchsearchsorted/src/cpu/searchsorted_cpu_wrapper.cpp:4-122 of the reference:
 *   - probe(): the three-way test of a candidate column (":4-39" eval): is `val` inside the gap
 *     (a[col], a[col+1]] (side left) / [a[col], a[col+1]) (side right), to its right, or to its left;
 *     the last column only has a left neighbour.
 *   - locate(): bisection over [0, ncol] that narrows with `left = mid` / `right = mid`, returning
 *     -1 when val lies before the first element and ncol-1 when it lies after the last (":41-80").
 *   - ss_oracle(): the serial double loop with row broadcasting, result = locate() + 1 as int64
 *     (":82-122").
 * Pinned against numpy.searchsorted and the compiled reference (oracle/_ref) in tests/test_searchsorted_oracle.py.
 * Build: `make -C oracle` -> oracle/_build/libsearchsorted_oracle.so
 */
#include <stdint.h>

static int probe(const float* row, int64_t col, int64_t ncol, float val, int side_left) {
  if (col == ncol - 1) return row[col] <= val ? 1 : -1;
  int lower_ok, upper_ok;
  if (side_left) {
    lower_ok = row[col] < val;
    upper_ok = row[col + 1] >= val;
  } else {
    lower_ok = row[col] <= val;
    upper_ok = row[col + 1] > val;
  }
  if (lower_ok && upper_ok) return 0;
  return lower_ok ? 1 : -1;
}

static int64_t locate(const float* row, int64_t ncol, float val, int side_left) {
  int64_t left = 0, right = ncol;
  while (right >= left) {
    const int64_t mid = left + (right - left) / 2;
    const int where = probe(row, mid, ncol, val, side_left);
    if (where == 0) return mid;
    if (where > 0) {
      if (mid == ncol - 1) return ncol - 1;
      left = mid;
    } else {
      if (mid == 0) return -1;
      right = mid;
    }
  }
  return -1;
}

void ss_oracle(const float* a, int64_t rows_a, int64_t na, const float* v, int64_t rows_v, int64_t nv, int64_t* res,
               int side_left) {
  const int64_t rows = rows_a > rows_v ? rows_a : rows_v;
  for (int64_t r = 0; r < rows; ++r) {
    const float* arow = a + (rows_a == 1 ? 0 : r) * na;
    const float* vrow = v + (rows_v == 1 ? 0 : r) * nv;
    for (int64_t c = 0; c < nv; ++c) res[r * nv + c] = locate(arow, na, vrow[c], side_left) + 1;
  }
}
