/*
 * TEST INFRASTRUCTURE ONLY -- see msda_oracle_impl.h for what this restates and why.
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against golden vectors
 * that tests/golden/gen_golden.py produced by running the reference's own
 * ms_deform_attn_core_pytorch (models/ops/functions/ms_deform_attn_func.py:41-61) and its
 * autograd gradients, on the input recipe of the reference's test (models/ops/test.py:21-36).
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared -fPIC) -> oracle/libmsda_oracle.so
 */
#include <math.h>
#include <stdint.h>

#define REAL double
#define SUFFIX _f64
#define FLOOR floor
#include "msda_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef FLOOR

#define REAL float
#define SUFFIX _f32
#define FLOOR floorf
#include "msda_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef FLOOR

int msda_oracle_abi_version(void) { return 1; }
