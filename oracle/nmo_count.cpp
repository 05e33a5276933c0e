// nmo_count.cpp — counters of the op-counting instantiation of the oracle (nmo_count.h); compiled WITHOUT the force-include.
#include <cstring>
namespace nmo_count {
enum { NOPS = 17 };
struct Counters { unsigned long long n[NOPS]; };
thread_local Counters tl = {};
}  // namespace nmo_count
extern "C" {
// ADD MUL DIV CMP EXP LOG LOG10 POW DPOW SQRT ATAN TAN COS SIN ASIN ACOS TANH of the calling thread since the last reset
void nmo_opcount_read(unsigned long long* out17, int reset) {
  std::memcpy(out17, nmo_count::tl.n, sizeof(nmo_count::tl.n));
  if (reset) std::memset(nmo_count::tl.n, 0, sizeof(nmo_count::tl.n));
}
}
