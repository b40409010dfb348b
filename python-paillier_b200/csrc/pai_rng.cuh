// pai_rng.cuh -- device-side obfuscators: r uniform in [1, n) for batched encryption.
//
// The reference draws every r from random.SystemRandom().randrange(1, n) (phe/paillier.py:141-143), one
// os.urandom call per element.  The batched engine expands ONE 256-bit seed from os.urandom with ChaCha20
// (D. J. Bernstein's original layout: 64-bit block counter, 64-bit nonce) on the device: element g, attempt t,
// block b uses counter (g << 12) | (t << 6) | b; an attempt takes bitlen(n) random bits and is accepted iff
// 1 <= r < n (rejection sampling: exactly uniform, >= 50 % acceptance).  Parity tests always inject r, so this
// never enters a bit-exactness claim; tests/test_rng_hostsim.py checks the keystream against an independent
// ChaCha20 and the range/determinism properties.
#pragma once
#include "pai_core.cuh"

namespace pai {

PAI_DEV uint32_t rotl32(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }

#define PAI_QR(a, b, c, d)                 \
  a += b; d ^= a; d = rotl32(d, 16);       \
  c += d; b ^= c; b = rotl32(b, 12);       \
  a += b; d ^= a; d = rotl32(d, 8);        \
  c += d; b ^= c; b = rotl32(b, 7);

PAI_DEV void chacha20_block(const uint32_t key[8], uint64_t counter, uint64_t nonce, uint32_t out[16]) {
  uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                    key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                    (uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)nonce, (uint32_t)(nonce >> 32)};
  uint32_t x[16];
  PAI_UNROLL
  for (int i = 0; i < 16; i++) x[i] = s[i];
  for (int r = 0; r < 10; r++) {
    PAI_QR(x[0], x[4], x[8], x[12]) PAI_QR(x[1], x[5], x[9], x[13]) PAI_QR(x[2], x[6], x[10], x[14]) PAI_QR(x[3], x[7], x[11], x[15])
    PAI_QR(x[0], x[5], x[10], x[15]) PAI_QR(x[1], x[6], x[11], x[12]) PAI_QR(x[2], x[7], x[8], x[13]) PAI_QR(x[3], x[4], x[9], x[14])
  }
  PAI_UNROLL
  for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}

// out_row[0..ln) = r uniform in [1, n);  n: ln limbs (padded), nbits = bit length of n
PAI_DEV void rng_fill_lt_n(const uint32_t key[8], uint64_t nonce, uint64_t g, const uint32_t* n, int ln, int nbits,
                           uint32_t* out_row) {
  const int used = (nbits + 31) / 32;                      // limbs that carry random bits
  const uint32_t topmask = (nbits & 31) ? ((1u << (nbits & 31)) - 1u) : 0xffffffffu;
  for (int i = used; i < ln; i++) out_row[i] = 0;
  for (int attempt = 0; attempt < 64; attempt++) {
    for (int b = 0; b * 16 < used; b++) {
      uint32_t blk[16];
      chacha20_block(key, (g << 12) | ((uint64_t)attempt << 6) | (uint64_t)b, nonce, blk);
      for (int i = 0; i < 16 && b * 16 + i < used; i++) out_row[b * 16 + i] = blk[i];
    }
    out_row[used - 1] &= topmask;
    // accept iff 1 <= r < n
    uint32_t nz = 0, bo = 0;
    for (int i = 0; i < used; i++) {
      nz |= out_row[i];
      uint64_t d = (uint64_t)out_row[i] - n[i] - bo;
      bo = (uint32_t)(d >> 63);
    }
    if (nz != 0 && bo == 1) return;
  }
  for (int i = 1; i < ln; i++) out_row[i] = 0;             // unreachable in practice (probability 2^-64)
  out_row[0] = 1;
}

}  // namespace pai
