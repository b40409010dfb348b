// TEST-ONLY host simulation of the device templates in pai_core.cuh (compiled with -DPAI_HOSTSIM).
// Never part of the product library.
#include <vector>
#include <cstring>
#include "../../python-paillier_b200/csrc/pai_core.cuh"
using namespace pai;

template <int NT>
static void run_mont(int sqr, const uint32_t* a, const uint32_t* b, const uint32_t* N, const uint32_t* ninv, uint32_t* out, int nthreads, int tid) {
  // simulate the interleaved layout with `nthreads` lanes, computing only lane `tid`
  const int Q = 2 * NT;
  std::vector<u4> A(Q * nthreads), B(Q * nthreads), O(Q * nthreads), Nc(Q), NI(2);
  memcpy(NI.data(), ninv, 32);
  for (int q = 0; q < Q; q++) {
    memcpy(&A[q * nthreads + tid], a + 4 * q, 16);
    memcpy(&B[q * nthreads + tid], b + 4 * q, 16);
    memcpy(&Nc[q], N + 4 * q, 16);
  }
  Opnd oa{A.data() + tid, nthreads}, ob{B.data() + tid, nthreads}, oo{O.data() + tid, nthreads}, on{Nc.data(), 1};
  Opnd oni{NI.data(), 1};
  if (sqr == 1) mont_sqr<NT>(oo, oa, on, oni); else mont_mul<NT>(oo, oa, ob, on, oni);
  for (int q = 0; q < Q; q++) memcpy(out + 4 * q, &O[q * nthreads + tid], 16);
}

extern "C" int sim_mont(int NT, int sqr, const uint32_t* a, const uint32_t* b, const uint32_t* N, const uint32_t* ninv, uint32_t* out) {
  switch (NT) {
    case 1: run_mont<1>(sqr, a, b, N, ninv, out, 3, 1); break;
    case 2: run_mont<2>(sqr, a, b, N, ninv, out, 3, 2); break;
    case 4: run_mont<4>(sqr, a, b, N, ninv, out, 5, 0); break;
    case 8: run_mont<8>(sqr, a, b, N, ninv, out, 2, 1); break;
    case 16: run_mont<16>(sqr, a, b, N, ninv, out, 4, 3); break;
    case 24: run_mont<24>(sqr, a, b, N, ninv, out, 1, 0); break;
    default: return -1;
  }
  return 0;
}
extern "C" void sim_tile(const uint32_t* a, const uint32_t* b, uint32_t* out16, uint32_t* lo8) {
  Acc A; acc_clear(A); tile_mac(A, a, b);
  uint32_t v[8]; acc_resolve_low(A, v); memcpy(out16, v, 32); acc_shift8(A); acc_resolve_low(A, v); memcpy(out16 + 8, v, 32);
  mul_lo8(lo8, a, b);
}

#include "../../python-paillier_b200/csrc/pai_digit.cuh"
template <int NTH>
static void run_digit(int sqr, const uint32_t* x, const uint32_t* y, const uint32_t* N, const uint32_t* ninv, const uint32_t* KL, uint32_t* out) {
  // x, y, out: 2*NTH tiles each as [d0 | d1]; simulate lane 1 of 3
  const int nthr = 3, tid = 1, Q = 2 * NTH;               // quads per digit
  std::vector<u4> X(2 * Q * nthr), Y(2 * Q * nthr), O(2 * Q * nthr), Nc(Q), NI(2), K(Q);
  for (int q = 0; q < 2 * Q; q++) { memcpy(&X[q * nthr + tid], x + 4 * q, 16); memcpy(&Y[q * nthr + tid], y + 4 * q, 16); }
  for (int q = 0; q < Q; q++) { memcpy(&Nc[q], N + 4 * q, 16); memcpy(&K[q], KL + 4 * q, 16); }
  memcpy(NI.data(), ninv, 32);
  Opnd bx{X.data() + tid, nthr}, by{Y.data() + tid, nthr}, bo{O.data() + tid, nthr};
  Opnd oN{Nc.data(), 1}, oNI{NI.data(), 1}, oK{K.data(), 1};
  std::vector<u4> N2v(Q), N3v(Q), TOPv(2);
  {
    const uint32_t* n32 = (const uint32_t*)Nc.data();
    uint32_t* a2 = (uint32_t*)N2v.data(); uint32_t* a3 = (uint32_t*)N3v.data(); uint32_t* tp = (uint32_t*)TOPv.data();
    uint64_t c = 0; for (int i = 0; i < 8 * NTH; i++) { c += 2ull * n32[i]; a2[i] = (uint32_t)c; c >>= 32; } tp[0] = (uint32_t)c;
    c = 0; for (int i = 0; i < 8 * NTH; i++) { c += 3ull * n32[i]; a3[i] = (uint32_t)c; c >>= 32; } tp[1] = (uint32_t)c;
    for (int i = 2; i < 8; i++) tp[i] = 0;
  }
  DigitEnv env;
  env.N = oN; env.NI = oNI; env.KL = oK; env.ONE = oN; env.ZERO = oN;
  env.N2 = Opnd{N2v.data(), 1}; env.N3 = Opnd{N3v.data(), 1}; env.TOPS = Opnd{TOPv.data(), 1};
  if (sqr) dsqr<NTH>(half_lo<NTH>(bo), half_hi<NTH>(bo), half_lo<NTH>(bx), half_hi<NTH>(bx), &env);
  else dmul<NTH>(half_lo<NTH>(bo), half_hi<NTH>(bo), half_lo<NTH>(bx), half_hi<NTH>(bx), half_lo<NTH>(by), half_hi<NTH>(by), &env);
  // result: Z0 in hi half, Z1 in lo half -> return as [Z0 | Z1]
  for (int q = 0; q < Q; q++) { memcpy(out + 4 * q, &O[(Q + q) * nthr + tid], 16); memcpy(out + 4 * (Q + q), &O[q * nthr + tid], 16); }
}
extern "C" int sim_digit(int NTH, int sqr, const uint32_t* x, const uint32_t* y, const uint32_t* N, const uint32_t* ninv, const uint32_t* KL, uint32_t* out) {
  switch (NTH) {
    case 1: run_digit<1>(sqr, x, y, N, ninv, KL, out); break;
    case 2: run_digit<2>(sqr, x, y, N, ninv, KL, out); break;
    case 3: run_digit<3>(sqr, x, y, N, ninv, KL, out); break;
    case 4: run_digit<4>(sqr, x, y, N, ninv, KL, out); break;
    case 8: run_digit<8>(sqr, x, y, N, ninv, KL, out); break;
    case 12: run_digit<12>(sqr, x, y, N, ninv, KL, out); break;
    default: return -1;
  }
  return 0;
}
