// ref_shim.cpp — the few symbols the generated translation of the reference needs around it (TEST INFRASTRUCTURE).
#include "ref_prelude.h"
int ref_math_mode = 0;
extern "C" void ref_set_math_mode(int m) { ref_math_mode = m; }
extern "C" int ref_get_math_mode(void) { return ref_math_mode; }
// never inlined, so that a constant exponent at the call site is not folded (see ref_prelude.h)
__attribute__((noinline)) float ref_powf(float x, float y) { return __builtin_powf(x, y); }
__attribute__((noinline)) double ref_pow(double x, double y) { return __builtin_pow(x, y); }
#ifdef REF_COVERAGE
unsigned char ref_cover[1000000];
extern "C" unsigned char* ref_cover_map(void) { return ref_cover; }
#endif
