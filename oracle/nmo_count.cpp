// nmo_count.cpp — counters of the op-counting instantiation of the oracle (nmo_count.h); compiled WITHOUT the force-include.
#include <cstring>
#include <vector>
namespace nmo_count {
struct TraceRec { int op; unsigned a, b, r; };
thread_local std::vector<TraceRec>* trace = nullptr;
enum { NOPS = 17 };
struct Counters { unsigned long long n[NOPS]; };
thread_local Counters tl = {};
}  // namespace nmo_count
extern "C" {
// value trace of the calling thread: start, run, then stop -> number of records, *out valid until the next start
void nmo_trace_start(void) {
  static thread_local std::vector<nmo_count::TraceRec> store;
  store.clear();
  nmo_count::trace = &store;
}
long nmo_trace_stop(const void** out) {
  long n = nmo_count::trace ? (long)nmo_count::trace->size() : 0;
  if (out) *out = nmo_count::trace ? (const void*)nmo_count::trace->data() : nullptr;
  nmo_count::trace = nullptr;
  return n;
}
// ADD MUL DIV CMP EXP LOG LOG10 POW DPOW SQRT ATAN TAN COS SIN ASIN ACOS TANH of the calling thread since the last reset
void nmo_opcount_read(unsigned long long* out17, int reset) {
  std::memcpy(out17, nmo_count::tl.n, sizeof(nmo_count::tl.n));
  if (reset) std::memset(nmo_count::tl.n, 0, sizeof(nmo_count::tl.n));
}
}
