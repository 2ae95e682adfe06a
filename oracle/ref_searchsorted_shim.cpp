// Build shim for the reference's OWN C++ searchsorted -- TEST INFRASTRUCTURE ONLY.
// The reference source is compiled from where it lies under /root/reference (nothing is copied); the
// only obstacle on torch 2.11 is `AT_DISPATCH_ALL_TYPES(a.type(), ...)` (a.type() no longer converts to
// a ScalarType), so torch's headers are included first and `type()` is then re-spelt for the one
// translation unit that follows.  REF_SEARCHSORTED_CPP is passed by oracle/Makefile.
#include <torch/extension.h>
#define type() scalar_type()
#include REF_SEARCHSORTED_CPP
