// nmp_io.cuh — per-column access to the HBM-resident data from inside the physics.
//
// The compact state is a structure of arrays: plane `slot` holds one word per active column, so the accesses
// of the 32 threads of a warp to one plane are a single unit-stride transaction.  The physics loads a state
// word as late as it is first needed and stores an output as soon as it is final, which keeps the live
// register set (and with it the spill traffic) of the 17k-instruction column program small.
#pragma once
#include "nmp_common.cuh"
#include "nmp_fields.h"

namespace nmp {

struct ColumnIO {
  const nmpf::StepParams& p;
  long long n;  // compact column
  int cell;     // grid cell of the column (forcing planes are in grid order)
  bool on;      // stores enabled (false for the padding threads of the last block and for rejected columns)
  // Address of plane `slot` = col + slot * np4: one 32x32->64-bit multiply-add per access (np * 4 < 2^32 because a
  // tile has fewer than 2^25 cells), instead of 64-bit index arithmetic followed by a scale by 4.
  char* col;  // &state[0 * np + n]; the plane stride in bytes, p.np4, is read from the parameter bank at each use
  __device__ ColumnIO(const nmpf::StepParams& p_, long long n_, bool on_)
      : p(p_), n(n_), cell(p_.cell[n_]), on(on_), col(reinterpret_cast<char*>(p_.state + n_)) {}
  __device__ float* at(int slot) const {
    // spelled as the instruction it should become: left to itself the compiler derives each plane address from the
    // previous one through chains of 64-bit adds that cost more than the multiply-add and lengthen live ranges
    unsigned long long a;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(a) : "r"(p.np4), "r"((unsigned)slot), "l"(col));
    float* q = reinterpret_cast<float*>(a);
    __builtin_assume(__isGlobal(q));
    return q;
  }
  // State, forcing and outputs are touched once per step, while the ~440-byte spill frame of the column program is
  // re-read all the time: the plane accesses bypass L1 allocation (ld.global.L1::no_allocate / st.global.cs) so that
  // L1 keeps the spill lines and the parameter tables (-7.5 % step time, profiles/r01_notes.md).
  static __device__ float ldstream(const float* q) {
    float v;
    asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(q));
    return v;
  }
  __device__ float forc(int f) const { return ldstream(p.forc[f] + cell); }
  // static inputs: compact copies (planes PLANE_STATIC0 + f), unit stride also after re-binning
  __device__ float stat(int f) const { return ldstream(at(nmpf::PLANE_STATIC0 + f)); }
  __device__ int stati(int f) const { return __float_as_int(stat(f)); }
  __device__ float ld(int slot) const { return ldstream(at(slot)); }
  __device__ int ldi(int slot) const { return __float_as_int(ldstream(at(slot))); }
  __device__ void st(int slot, float v) const { if (on) __stcs(at(slot), v); }
  __device__ void sti(int slot, int v) const { if (on) __stcs(at(slot), __int_as_float(v)); }
  // accumulator += v.  Production build: one fire-and-forget RED.ADD.F32 (no load latency in the dependency chain;
  // a column has a single writer, so the sum is the same round-to-nearest add).  Parity build: load-add-store,
  // because red.add.f32 flushes subnormals and the oracle does not.
  __device__ void acc(int slot, float v) const {
    if (!on) return;
#if NMP_FASTMATH
    atomicAdd(at(slot), v);
#else
    *at(slot) = *at(slot) + v;
#endif
  }
  // L2 prefetch of a plane element that is loaded late in the column program
  __device__ void prefetch(int slot) const { asm volatile("prefetch.global.L2 [%0];" ::"l"(at(slot))); }
};

}  // namespace nmp
