// kgx_hosttest.cpp -- CPU build of the host/device-portable pieces of the engine (currently the
// safegcd modular inverse) so they can be unit-tested without a GPU (tests/test_modinv_host.py).
// Not part of the product data path.
#include "kgx_modinv.h"
extern "C" void kgx_host_modinv(uint64_t out[4], const uint64_t in[4]) { kgx::modinv256(out, in); }
