// pai_coop.cuh -- one WARP per ciphertext: the low-latency form of the hot path for small batches and batch tails.
//
// The throughput kernels (pai_cta.cuh) give every ciphertext one thread; a lone r^n mod n^2 then takes the full
// ~0.2 s of a 2048-bit ladder whatever the batch size, where the reference's GMP call (phe/util.py:50) takes 9 ms.
// Here the 32 lanes of a warp share one number: lane t holds limbs [t*K, (t+1)*K) in registers, a Montgomery
// product is the word-serial CIOS loop with the multiplier limb broadcast by shuffle, accumulators stay in
// carry-save form (64-bit per limb position, no carry chains inside the loop), and carries / borrows between lanes
// are resolved once per product with ballot-terminated ripple rounds.  Window table (fixed 4-bit windows, always multiplied: the exponent
// may be secret) in shared memory.  Same results as the thread-per-ciphertext kernels, bit for bit.
//
// Written over "lane arrays" x[WL][..]: WL = 1 on the GPU (plain registers, ln is always 0) and WL = 32 in the CPU
// simulation build of the tests, where one call walks all lanes of the warp in lockstep statement by statement.
#pragma once
#include "pai_core.cuh"

namespace pai {

#ifdef PAI_HOSTSIM
constexpr int WL = 32;
#define PAI_LANE(ln) (ln)
#else
constexpr int WL = 1;
#define PAI_LANE(ln) ((int)(threadIdx.x & 31u))
#endif
#define PAI_EACH_LANE for (int ln = 0; ln < WL; ln++)
constexpr int COOP_W = 4;                 // window bits
constexpr int COOP_WARPS = 1;             // warps per CTA: one, so that all control flow around the shuffles is
                                          // block-uniform and ptxas emits them without divergence guards

// ---- warp primitives over lane arrays -----------------------------------------------------------------------
template <class T>
PAI_DEV void w_bcast(T (&dst)[WL], const T (&src)[WL], int src_lane) {
#ifdef PAI_HOSTSIM
  T v = src[src_lane];
  PAI_EACH_LANE dst[ln] = v;
#else
  dst[0] = __shfl_sync(0xffffffffu, src[0], src_lane);
#endif
}
// dst[lane] = src[lane + 1], 0 in lane 31
template <class T>
PAI_DEV void w_down1(T (&dst)[WL], const T (&src)[WL]) {
#ifdef PAI_HOSTSIM
  for (int ln = 0; ln < 31; ln++) dst[ln] = src[ln + 1];
  dst[31] = 0;
#else
  T v = __shfl_down_sync(0xffffffffu, src[0], 1);
  dst[0] = (threadIdx.x & 31u) == 31u ? (T)0 : v;
#endif
}
// dst[lane] = src[lane - 1], 0 in lane 0
template <class T>
PAI_DEV void w_up1(T (&dst)[WL], const T (&src)[WL]) {
#ifdef PAI_HOSTSIM
  for (int ln = 31; ln > 0; ln--) dst[ln] = src[ln - 1];
  dst[0] = 0;
#else
  T v = __shfl_up_sync(0xffffffffu, src[0], 1);
  dst[0] = (threadIdx.x & 31u) == 0u ? (T)0 : v;
#endif
}
PAI_DEV uint32_t w_ballot(const bool (&p)[WL]) {
#ifdef PAI_HOSTSIM
  uint32_t m = 0;
  PAI_EACH_LANE m |= (p[ln] ? 1u : 0u) << ln;
  return m;
#else
  return __ballot_sync(0xffffffffu, p[0]);
#endif
}
PAI_DEV void w_sync() {
#ifndef PAI_HOSTSIM
  __syncwarp();
#endif
}

// ---- number <-> memory ----------------------------------------------------------------------------------------
// lane t takes limbs [t*K, t*K+K) of a row of `limbs` limbs starting at limb `off` (zero beyond the row)
template <int K>
PAI_DEV void w_load(uint32_t (&x)[WL][K], const uint32_t* row, int limbs, int off = 0) {
  PAI_EACH_LANE {
    const int lane = PAI_LANE(ln);
    PAI_UNROLL
    for (int k = 0; k < K; k++) { int i = off + lane * K + k; x[ln][k] = i < limbs ? row[i] : 0u; }
  }
}
template <int K>
PAI_DEV void w_store(uint32_t* row, int limbs, const uint32_t (&x)[WL][K]) {
  PAI_EACH_LANE {
    const int lane = PAI_LANE(ln);
    PAI_UNROLL
    for (int k = 0; k < K; k++) { int i = lane * K + k; if (i < limbs) row[i] = x[ln][k]; }
  }
}
// shared-memory table entry e of this warp: word (e*K + k)*32 + lane
template <int K>
PAI_DEV void w_tbl_put(uint32_t* tbl, int e, const uint32_t (&x)[WL][K]) {
  PAI_EACH_LANE {
    PAI_UNROLL
    for (int k = 0; k < K; k++) tbl[(e * K + k) * 32 + PAI_LANE(ln)] = x[ln][k];
  }
}
template <int K>
PAI_DEV void w_tbl_get(uint32_t (&x)[WL][K], const uint32_t* tbl, int e) {
  PAI_EACH_LANE {
    PAI_UNROLL
    for (int k = 0; k < K; k++) x[ln][k] = tbl[(e * K + k) * 32 + PAI_LANE(ln)];
  }
}
template <int K>
PAI_DEV void w_set_small(uint32_t (&x)[WL][K], uint32_t v) {
  PAI_EACH_LANE {
    PAI_UNROLL
    for (int k = 0; k < K; k++) x[ln][k] = (PAI_LANE(ln) == 0 && k == 0) ? v : 0u;
  }
}

// ---- carries between lanes ------------------------------------------------------------------------------------
// x += (carry words sp[lane] entering lane + 1); returns the carry out of lane 31 (0 or small)
template <int K>
PAI_DEV uint32_t w_resolve_carries(uint32_t (&x)[WL][K], uint64_t (&sp)[WL]) {
  uint32_t top = 0;
  for (;;) {
    uint64_t cin[WL];
    uint64_t last[WL];
    w_bcast(last, sp, 31);
    top += (uint32_t)last[0];
    w_up1(cin, sp);
    bool any[WL];
    PAI_EACH_LANE {
      uint64_t c = cin[ln];
      PAI_UNROLL
      for (int k = 0; k < K; k++) { uint64_t t = (uint64_t)x[ln][k] + c; x[ln][k] = (uint32_t)t; c = t >> 32; }
      sp[ln] = c;
      any[ln] = c != 0;
    }
    if (!w_ballot(any)) break;
  }
  return top;
}
// x -= n (whole warp), the borrow out of lane 31 is dropped (the caller knows the result is non-negative mod 2^(32*32K))
template <int K>
PAI_DEV void w_sub(uint32_t (&x)[WL][K], const uint32_t (&n)[WL][K]) {
  uint32_t bo[WL];
  PAI_EACH_LANE {
    uint32_t b = 0;
    PAI_UNROLL
    for (int k = 0; k < K; k++) {
      uint64_t d = (uint64_t)x[ln][k] - n[ln][k] - b;
      x[ln][k] = (uint32_t)d; b = (uint32_t)(d >> 63);
    }
    bo[ln] = b;
  }
  for (;;) {
    uint32_t bin[WL];
    w_up1(bin, bo);
    bool any[WL];
    PAI_EACH_LANE {
      uint32_t b = bin[ln];
      PAI_UNROLL
      for (int k = 0; k < K; k++) { uint64_t d = (uint64_t)x[ln][k] - b; x[ln][k] = (uint32_t)d; b = (uint32_t)(d >> 63); }
      bo[ln] = b;
      any[ln] = b != 0 && PAI_LANE(ln) != 31;
    }
    if (!w_ballot(any)) break;
  }
}
// x >= n ?
template <int K>
PAI_DEV bool w_geq(const uint32_t (&x)[WL][K], const uint32_t (&n)[WL][K]) {
  bool gt[WL], lt[WL];
  PAI_EACH_LANE {
    bool g = false, l = false;
    PAI_UNROLL
    for (int k = K - 1; k >= 0; k--) {
      if (!g && !l) { g = x[ln][k] > n[ln][k]; l = x[ln][k] < n[ln][k]; }
    }
    gt[ln] = g; lt[ln] = l;
  }
  return w_ballot(gt) >= w_ballot(lt);
}
template <int K>
PAI_DEV void w_cond_sub(uint32_t (&x)[WL][K], uint32_t top, const uint32_t (&n)[WL][K]) {
  if (top || w_geq<K>(x, n)) w_sub<K>(x, n);
}
// x = (x + y) mod n for x, y < n
template <int K>
PAI_DEV void w_add_mod(uint32_t (&x)[WL][K], const uint32_t (&y)[WL][K], const uint32_t (&n)[WL][K]) {
  uint64_t sp[WL];
  PAI_EACH_LANE {
    uint64_t c = 0;
    PAI_UNROLL
    for (int k = 0; k < K; k++) { uint64_t t = (uint64_t)x[ln][k] + y[ln][k] + c; x[ln][k] = (uint32_t)t; c = t >> 32; }
    sp[ln] = c;
  }
  uint32_t top = w_resolve_carries<K>(x, sp);
  w_cond_sub<K>(x, top, n);
}

// ---- Montgomery product: r = a * b / 2^(32*32K) mod n, a < 2^(32*32K), b < n, r < n ----------------------------
// Word-serial CIOS in carry-save form: position k of a lane is a 64-bit accumulator that is < 2^33 whenever a row of
// products is added (so a[k]*b + acc[k] cannot overflow: one IMAD.WIDE per product, all K independent), and the
// high halves move up one position when the number is shifted down a limb.  The only values that cross lanes per
// limb are the multiplier limb, the quotient limb (lane 0's low word) and the low word each lane hands down.
template <int K>
PAI_DEV void w_mont_mul(uint32_t (&r)[WL][K], const uint32_t (&a)[WL][K], const uint32_t (&b)[WL][K],
                        const uint32_t (&n)[WL][K], uint32_t n0inv) {
  uint64_t acc[WL][K];
  uint64_t sp[WL];                         // carry word at the position just above the lane's K limbs (small)
  uint32_t pend[WL];                       // low word handed down by the lane above, one iteration late (K >= 2): the
                                           // SHFL.DOWN then overlaps the next limb's products and quotient broadcast
  uint32_t bj[WL], bsel[WL];
  PAI_EACH_LANE {
    sp[ln] = 0; pend[ln] = 0;
    PAI_UNROLL
    for (int k = 0; k < K; k++) acc[ln][k] = 0;
    bsel[ln] = b[ln][0];
  }
  w_bcast(bj, bsel, 0);
  for (int jt = 0; jt < 32; jt++) {
    PAI_UNROLL
    for (int jk = 0; jk < K; jk++) {
      uint32_t bnext[WL], a0[WL], q[WL], low[WL], nl[WL];
      PAI_EACH_LANE bsel[ln] = b[ln][(jk + 1) % K];
      w_bcast(bnext, bsel, (jt + (jk + 1 == K ? 1 : 0)) & 31);       // multiplier limb of the next iteration
      PAI_EACH_LANE {
        PAI_UNROLL
        for (int k = 0; k < K; k++) acc[ln][k] += (uint64_t)a[ln][k] * bj[ln];
        a0[ln] = (uint32_t)acc[ln][0];
      }
      w_bcast(q, a0, 0);
      PAI_EACH_LANE {
        const uint32_t qq = q[ln] * n0inv;
        if (K >= 2) acc[ln][K - 1] += pend[ln];
        // high halves up one position (the top one into the lane's carry word), then the quotient row
        sp[ln] += acc[ln][K - 1] >> 32;
        PAI_UNROLL
        for (int k = K - 1; k >= 1; k--) acc[ln][k] = (uint64_t)(uint32_t)acc[ln][k] + (acc[ln][k - 1] >> 32);
        acc[ln][0] = (uint32_t)acc[ln][0];
        PAI_UNROLL
        for (int k = 0; k < K; k++) acc[ln][k] += (uint64_t)n[ln][k] * qq;
        low[ln] = (uint32_t)acc[ln][0];
      }
      // lane 0's low word is zero now: every position moves down one limb, high halves up one position
      w_down1(nl, low);
      PAI_EACH_LANE {
        PAI_UNROLL
        for (int k = 0; k + 1 < K; k++) acc[ln][k] = (uint64_t)(uint32_t)acc[ln][k + 1] + (acc[ln][k] >> 32);
        uint64_t v = (acc[ln][K - 1] >> 32) + sp[ln] + (K >= 2 ? 0u : nl[ln]);
        acc[ln][K - 1] = (uint32_t)v;
        sp[ln] = v >> 32;
        pend[ln] = nl[ln];
        bj[ln] = bnext[ln];
      }
    }
  }
  // positions are < 2^33: normalise inside the lane, then across lanes
  uint32_t out[WL][K];
  PAI_EACH_LANE {
    if (K >= 2) acc[ln][K - 1] += pend[ln];
    uint64_t c = 0;
    PAI_UNROLL
    for (int k = 0; k < K; k++) { uint64_t t = acc[ln][k] + c; out[ln][k] = (uint32_t)t; c = t >> 32; }
    sp[ln] += c;
  }
  uint32_t top = w_resolve_carries<K>(out, sp);
  w_cond_sub<K>(out, top, n);
  PAI_EACH_LANE {
    PAI_UNROLL
    for (int k = 0; k < K; k++) r[ln][k] = out[ln][k];
  }
}

// ---- programs ---------------------------------------------------------------------------------------------------
// constants of one modulus for this layout: [ N | RR = R^2 mod N | RRR = R^3 mod N ], R = 2^(32*32K), 32K limbs each
template <int K>
struct CoopC {
  uint32_t n[WL][K], rr[WL][K];
  uint32_t n0inv;                          // -N^-1 mod 2^32
};
template <int K>
PAI_DEV void coop_bind(CoopC<K>& c, const uint32_t* blob, uint32_t n0inv) {
  w_load<K>(c.n, blob, 32 * K);
  w_load<K>(c.rr, blob + 32 * K, 32 * K);
  c.n0inv = n0inv;
}

// x <- x^e (Montgomery form in, Montgomery form out); e: nwin windows of COOP_W bits, shared by the batch
template <int K>
PAI_DEV void coop_pow(uint32_t (&x)[WL][K], const CoopC<K>& c, const uint32_t* e, int e_limbs, int nwin, uint32_t* tbl) {
  uint32_t t[WL][K], one[WL][K];
  w_set_small<K>(one, 1u);
  w_mont_mul<K>(t, c.rr, one, c.n, c.n0inv);            // R mod N
  w_tbl_put<K>(tbl, 0, t);
  w_tbl_put<K>(tbl, 1, x);
  PAI_EACH_LANE { PAI_UNROLL for (int k = 0; k < K; k++) t[ln][k] = x[ln][k]; }
  for (int i = 2; i < (1 << COOP_W); i++) {
    w_mont_mul<K>(t, t, x, c.n, c.n0inv);
    w_tbl_put<K>(tbl, i, t);
  }
  w_sync();
  w_tbl_get<K>(x, tbl, (int)exp_digit(e, e_limbs, (nwin - 1) * COOP_W, COOP_W));
  for (int wi = nwin - 2; wi >= 0; wi--) {
    for (int s = 0; s < COOP_W; s++) w_mont_mul<K>(x, x, x, c.n, c.n0inv);
    w_tbl_get<K>(t, tbl, (int)exp_digit(e, e_limbs, wi * COOP_W, COOP_W));
    w_mont_mul<K>(x, x, t, c.n, c.n0inv);
  }
  w_sync();
}

// base row of up to 2 * 32K limbs -> Montgomery form (reduced)
template <int K>
PAI_DEV void coop_to_mont(uint32_t (&x)[WL][K], const CoopC<K>& c, const uint32_t* blob, const uint32_t* row, int limbs) {
  uint32_t lo[WL][K];
  w_load<K>(lo, row, limbs, 0);
  w_mont_mul<K>(x, lo, c.rr, c.n, c.n0inv);
  if (limbs > 32 * K) {
    uint32_t hi[WL][K], rrr[WL][K];
    w_load<K>(hi, row, limbs, 32 * K);
    w_load<K>(rrr, blob + 2 * 32 * K, 32 * K);
    w_mont_mul<K>(hi, hi, rrr, c.n, c.n0inv);
    w_add_mod<K>(x, hi, c.n);
  }
}

// out[g] = base[g] ^ e mod N                                                   (util.powmod, phe/util.py:38-50)
template <int K>
PAI_DEV void coop_powmod(const uint32_t* blob, uint32_t n0inv, const uint32_t* base, int base_limbs, const uint32_t* e,
                         int e_limbs, int nwin, uint32_t* out, int out_limbs, uint32_t* tbl) {
  CoopC<K> c;
  coop_bind<K>(c, blob, n0inv);
  uint32_t x[WL][K], one[WL][K];
  coop_to_mont<K>(x, c, blob, base, base_limbs);
  if (nwin > 0) coop_pow<K>(x, c, e, e_limbs, nwin, tbl);
  w_set_small<K>(one, 1u);
  if (nwin > 0) w_mont_mul<K>(x, x, one, c.n, c.n0inv);
  else { w_mont_mul<K>(x, c.rr, one, c.n, c.n0inv); w_mont_mul<K>(x, x, one, c.n, c.n0inv); }   // e = 0: 1 mod N
  w_store<K>(out, out_limbs, x);
}

// c = (1 + n*m) * r^n mod n^2                                                  (raw_encrypt, phe/paillier.py:102-139)
template <int K>
PAI_DEV void coop_encrypt(const uint32_t* blob, uint32_t n0inv, const uint32_t* nrow, int n_limbs, const uint32_t* e,
                          int nwin, const uint32_t* m, const uint32_t* r, uint32_t* out, int out_limbs, uint32_t* tbl) {
  CoopC<K> c;
  coop_bind<K>(c, blob, n0inv);
  uint32_t x[WL][K], y[WL][K], one[WL][K];
  coop_to_mont<K>(x, c, blob, r, n_limbs);
  coop_pow<K>(x, c, e, n_limbs, nwin, tbl);                       // r^n * R
  // (1 + n*m) * R = mont(m, n*R^2) + R
  w_load<K>(y, nrow, n_limbs);
  w_mont_mul<K>(y, y, c.rr, c.n, c.n0inv);                        // n*R
  w_mont_mul<K>(y, y, c.rr, c.n, c.n0inv);                        // n*R^2
  uint32_t mm[WL][K];
  w_load<K>(mm, m, n_limbs);
  w_mont_mul<K>(y, mm, y, c.n, c.n0inv);                          // m*n*R   (m < 2^(32*n_limbs) <= R)
  w_set_small<K>(one, 1u);
  w_mont_mul<K>(mm, c.rr, one, c.n, c.n0inv);                     // R mod N
  w_add_mod<K>(y, mm, c.n);
  w_mont_mul<K>(x, x, y, c.n, c.n0inv);                           // (1+nm) r^n R
  w_mont_mul<K>(x, x, one, c.n, c.n0inv);
  w_store<K>(out, out_limbs, x);
}

}  // namespace pai
