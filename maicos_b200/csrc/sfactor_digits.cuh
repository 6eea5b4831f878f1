// Digit arithmetic shared by the table kernel (Z image) and the tensor-core contraction (sfactor_tc.cuh).
#pragma once
#include <cstdint>

namespace mb200 {

constexpr int TC_KA = 16;          // atoms per MMA K-step (K' = 32 bytes = 16 atoms x {re, im})
constexpr int TC_S = 5;            // digits per component

// byte offset of element (row, k) in the UMMA canonical K-major no-swizzle layout: 8 x 16 B core matrices,
// LBO = 128 B between the two K chunks, SBO = 256 B between 8-row groups
__device__ __host__ __forceinline__ int tc_canon(int row, int k) { return (row >> 3) * 256 + (k >> 4) * 128 + (row & 7) * 16 + (k & 15); }

// Digits: y = x + TC_MAGIC lies in [2^14, 2^15) (ulp 2^-38), so the low 40 mantissa bits of y are
// U = rint(x * 2^38) + 128 * (256^4 + 256^3 + 256^2 + 256 + 1); byte i of U xor 0x80 is the balanced digit of
// weight 256^i.  Returned: lo = digits 5, 4, 3, 2 (bytes 0..3), hi = digit 1 (byte 0), already re-centred.
#define TC_MAGIC (16384.0 + 551911719040.0 / 274877906944.0)
__device__ __forceinline__ void tc_digits(double x, unsigned &lo, unsigned &hi) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x + TC_MAGIC);
    lo = (unsigned)b ^ 0x80808080u;
    hi = ((unsigned)(b >> 32) ^ 0x80u) & 0xffu;
}
// (re, im) digit pair of slice s (1 = most significant) as a 16-bit value, re in the low byte
__device__ __forceinline__ unsigned tc_pair(int s, unsigned rlo, unsigned rhi, unsigned ilo, unsigned ihi) {
    switch (s) {
        case 1: return __byte_perm(rhi, ihi, 0x0040);
        case 2: return __byte_perm(rlo, ilo, 0x0073);
        case 3: return __byte_perm(rlo, ilo, 0x0062);
        case 4: return __byte_perm(rlo, ilo, 0x0051);
        default: return __byte_perm(rlo, ilo, 0x0040);
    }
}


}  // namespace mb200
