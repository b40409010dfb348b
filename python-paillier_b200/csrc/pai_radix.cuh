// pai_radix.cuh -- limb rows <-> decimal text, the radix conversion behind the reference's JSON wire format
// (docs/serialisation.rst:24-42: ciphertexts travel as str(int); phe/command_line.py:120-131 the same for "v").
// CPython needs ~30 us for str() and ~18 us for int() of one 4096-bit ciphertext, several times the cost of
// encrypting it on this engine, so the conversion runs on the device: one thread per number, the number held in
// shared memory (limb l of thread t at sm[l * nthreads + t]), one 10^9 chunk (nine digits) per pass.
// Text rows have a fixed width of pai_decimal_width(limbs) = 9 * chunks characters, right aligned, '0' padded.
#pragma once
#include "pai_core.cuh"

namespace pai {

PAI_HD int radix_chunks(int limbs) { return (int)(((long)limbs * 32 * 30103 + 99999) / 100000 + 8) / 9; }

// text[0..9*chunks) <- decimal digits of row[0..L)
PAI_DEV void radix_to_decimal(uint32_t* sm, int tid, int nthr, const uint32_t* row, int L, uint8_t* text, int chunks) {
  int top = -1;
  for (int i = 0; i < L; i++) { uint32_t v = row[i]; sm[i * nthr + tid] = v; if (v) top = i; }
  for (int c = chunks - 1; c >= 0; c--) {
    uint64_t rem = 0;
    for (int i = top; i >= 0; i--) {
      uint64_t cur = (rem << 32) | sm[i * nthr + tid];
      uint64_t q = cur / 1000000000ull;
      rem = cur - q * 1000000000ull;
      sm[i * nthr + tid] = (uint32_t)q;
    }
    while (top >= 0 && sm[top * nthr + tid] == 0) top--;
    uint32_t r = (uint32_t)rem;
    uint8_t* o = text + c * 9;
    for (int d = 8; d >= 0; d--) { uint32_t q = r / 10u; o[d] = (uint8_t)('0' + (r - q * 10u)); r = q; }
  }
}

// row[0..L) <- value of the decimal text; returns 0, 1 = a character that is not a digit, 2 = does not fit
PAI_DEV int radix_from_decimal(uint32_t* sm, int tid, int nthr, const uint8_t* text, int width, uint32_t* row, int L) {
  int top = -1, status = 0;
  int pos = 0;
  const int head = width % 9;                              // a first, shorter chunk when width is not a multiple of 9
  while (pos < width) {
    const int len = (pos == 0 && head) ? head : 9;
    uint32_t chunk = 0, scale = 1;
    for (int d = 0; d < len; d++) {
      uint32_t ch = text[pos + d];
      if (ch < '0' || ch > '9') { status = 1; ch = '0'; }
      chunk = chunk * 10u + (ch - '0');
      scale *= 10u;
    }
    pos += len;
    uint64_t carry = chunk;
    for (int i = 0; i <= top; i++) {
      uint64_t v = (uint64_t)sm[i * nthr + tid] * scale + carry;
      sm[i * nthr + tid] = (uint32_t)v;
      carry = v >> 32;
    }
    if (carry) {
      if (top + 1 < L) { top++; sm[top * nthr + tid] = (uint32_t)carry; }
      else if (!status) status = 2;
    }
  }
  for (int i = 0; i < L; i++) row[i] = (i <= top && status == 0) ? sm[i * nthr + tid] : 0u;
  return status;
}

}  // namespace pai
