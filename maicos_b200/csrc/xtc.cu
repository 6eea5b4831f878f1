// GROMACS .xtc frames (compressed coordinates) decoded on the host: the trajectory-ingest side of the profile / S(q) paths
// ("hand-written TRR/GRO readers, optional XTC decode").  MDAnalysis reads these files through
// the xdrfile library (third party, not in /root/reference); this is a restatement of the published xtc3 format:
//
//   frame  := magic 1995 | natoms | step | time f32 | box 9 x f32 | natoms | coordinates        (XDR: big-endian words)
//   natoms <= 9: 3 natoms raw floats
//   else: precision f32 | minint[3] | maxint[3] | smallidx | byte count | packed bits, padded to a multiple of four
//
//   packed bits (most significant bit first): every atom either as one mixed-radix number in the box of the whole frame
//   (maxint - minint + 1 per axis), or -- inside a run announced by 1 flag bit + 5 bits -- as a small offset to the atom
//   before it in a cube of magicints[smallidx]^3; the first two atoms of a run are swapped back (water: H O H -> O H H), and
//   the run length's remainder mod 3 moves smallidx up or down for the next atoms.
//
// No device work: plain C++ in a .cu file so that the one Makefile rule builds it.
#include "common.cuh"

using mb200::fail;

namespace {

const int kMagicInts[] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512,
                          645, 812, 1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007,
                          32768, 41285, 52015, 65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561,
                          832255, 1048576, 1321122, 1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983,
                          13316085, 16777216};
constexpr int kFirstIdx = 9;
constexpr int kLastIdx = (int)(sizeof(kMagicInts) / sizeof(kMagicInts[0]));

inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]; }
inline float be_f32(const uint8_t *p) {
    const uint32_t u = be32(p);
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

struct BitReader {  // most significant bit first
    const uint8_t *p;
    size_t n, pos = 0;  // pos in bits
    bool overrun = false;
    uint32_t bits(int k) {  // k <= 32
        uint32_t v = 0;
        while (k > 0) {
            const size_t byte = pos >> 3;
            if (byte >= n) { overrun = true; return 0; }
            const int avail = 8 - (int)(pos & 7);
            const int take = k < avail ? k : avail;
            const uint32_t chunk = ((uint32_t)p[byte] >> (avail - take)) & ((1u << take) - 1u);
            v = (v << take) | chunk;
            pos += (size_t)take;
            k -= take;
        }
        return v;
    }
};

int bit_length_u32(uint32_t size) {  // xdrfile's sizeofint: bits needed for values 0 .. size
    int bits = 0;
    uint32_t num = 1;
    while (size >= num && bits < 32) { ++bits; num <<= 1; }
    return bits;
}

int bit_length_product(const uint32_t sizes[3]) {  // xdrfile's sizeofints: bit length of sizes[0] * sizes[1] * sizes[2]
    unsigned __int128 prod = (unsigned __int128)sizes[0] * sizes[1] * sizes[2];
    int bits = 0;
    while (prod) { ++bits; prod >>= 1; }
    return bits;
}

// three numbers packed as one mixed-radix integer of `nbits` bits; the stream holds it least significant BYTE first, the
// last (partial) byte most significant
void unpack3(BitReader &br, int nbits, const uint32_t sizes[3], int32_t out[3]) {
    unsigned __int128 v = 0;
    int shift = 0;
    while (nbits > 8) { v |= (unsigned __int128)br.bits(8) << shift; shift += 8; nbits -= 8; }
    if (nbits > 0) v |= (unsigned __int128)br.bits(nbits) << shift;
    out[2] = (int32_t)(uint32_t)(v % sizes[2]); v /= sizes[2];
    out[1] = (int32_t)(uint32_t)(v % sizes[1]); v /= sizes[1];
    out[0] = (int32_t)(uint32_t)v;
}

}  // namespace

extern "C" int mb200_xtc_frame(const void *data, size_t size, size_t offset, int64_t *natoms_out, int32_t *step, float *time,
                               float box[9], float *positions, double scale, size_t *next_offset) {
    const uint8_t *d = reinterpret_cast<const uint8_t *>(data);
    if (!d || offset + 56 > size) return fail(MB200_EINVAL, "truncated XTC frame header at byte %zu", offset);
    const uint8_t *p = d + offset;
    const uint32_t magic = be32(p);
    if (magic != 1995u) return fail(MB200_EINVAL, "not an XTC frame at byte %zu (magic %u; large-system XTC, magic 2023, is not supported)", offset, magic);
    const int64_t natoms = (int32_t)be32(p + 4);
    if (natoms < 0) return fail(MB200_EINVAL, "negative atom count in the XTC frame at byte %zu", offset);
    if (natoms_out) *natoms_out = natoms;
    if (step) *step = (int32_t)be32(p + 8);
    if (time) *time = be_f32(p + 12);
    if (box) for (int i = 0; i < 9; ++i) box[i] = be_f32(p + 16 + 4 * i);
    if ((int64_t)(int32_t)be32(p + 52) != natoms) return fail(MB200_EINVAL, "inconsistent atom counts in the XTC frame at byte %zu", offset);
    size_t off = offset + 56;
    const float sc = (float)scale;
    if (natoms <= 9) {  // uncompressed
        if (off + (size_t)natoms * 12 > size) return fail(MB200_EINVAL, "truncated XTC frame at byte %zu", offset);
        if (positions) for (int64_t i = 0; i < natoms * 3; ++i) positions[i] = be_f32(d + off + 4 * i) * sc;
        if (next_offset) *next_offset = off + (size_t)natoms * 12;
        return MB200_OK;
    }
    if (off + 36 > size) return fail(MB200_EINVAL, "truncated XTC frame at byte %zu", offset);
    const float precision = be_f32(d + off);
    int32_t minint[3], maxint[3];
    for (int i = 0; i < 3; ++i) { minint[i] = (int32_t)be32(d + off + 4 + 4 * i); maxint[i] = (int32_t)be32(d + off + 16 + 4 * i); }
    int smallidx = (int32_t)be32(d + off + 28);
    const size_t nbytes = be32(d + off + 32);
    off += 36;
    const size_t padded = (nbytes + 3) / 4 * 4;
    if (off + padded > size) return fail(MB200_EINVAL, "truncated XTC frame at byte %zu", offset);
    if (next_offset) *next_offset = off + padded;
    if (!positions) return MB200_OK;  // scanning pass: headers only
    if (!(precision > 0.f) || smallidx < kFirstIdx || smallidx >= kLastIdx)
        return fail(MB200_EINVAL, "corrupt XTC frame at byte %zu (precision %g, smallidx %d)", offset, (double)precision, smallidx);

    uint32_t sizeint[3];
    int bitsizeint[3] = {0, 0, 0}, bitsize;
    for (int i = 0; i < 3; ++i) {
        if (maxint[i] < minint[i]) return fail(MB200_EINVAL, "corrupt XTC frame at byte %zu (empty coordinate range)", offset);
        sizeint[i] = (uint32_t)((int64_t)maxint[i] - (int64_t)minint[i] + 1);
    }
    if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffffu) {  // large ranges: one plain bit field per coordinate
        for (int i = 0; i < 3; ++i) bitsizeint[i] = bit_length_u32(sizeint[i]);
        bitsize = 0;
    } else {
        bitsize = bit_length_product(sizeint);
    }
    int smaller = kMagicInts[std::max(kFirstIdx, smallidx - 1)] / 2;
    int smallnum = kMagicInts[smallidx] / 2;
    uint32_t sizesmall[3] = {(uint32_t)kMagicInts[smallidx], (uint32_t)kMagicInts[smallidx], (uint32_t)kMagicInts[smallidx]};
    const float inv = 1.0f / precision;  // xdrfile: int * (1 / precision) in float32; the unit factor is applied afterwards
    BitReader br{d + off, nbytes};
    int64_t i = 0;
    int run = 0;
    float *out = positions;
    while (i < natoms) {
        int32_t cur[3], prev[3];
        if (bitsize == 0) {
            for (int k = 0; k < 3; ++k) cur[k] = (int32_t)br.bits(bitsizeint[k]);
        } else {
            unpack3(br, bitsize, sizeint, cur);
        }
        ++i;
        for (int k = 0; k < 3; ++k) { cur[k] += minint[k]; prev[k] = cur[k]; }
        const uint32_t flag = br.bits(1);
        int is_smaller = 0;
        if (flag) {
            run = (int)br.bits(5);
            is_smaller = run % 3;
            run -= is_smaller;
            --is_smaller;
        }
        if (br.overrun) return fail(MB200_EINVAL, "corrupt XTC frame at byte %zu (bit stream too short)", offset);
        if (run > 0) {
            if (i + run / 3 > natoms) return fail(MB200_EINVAL, "corrupt XTC frame at byte %zu (run past the last atom)", offset);
            for (int k = 0; k < run; k += 3) {
                unpack3(br, smallidx, sizesmall, cur);
                ++i;
                for (int c = 0; c < 3; ++c) cur[c] += prev[c] - smallnum;
                if (k == 0) {  // the first two atoms of a run were swapped by the writer (water: the oxygen goes first)
                    for (int c = 0; c < 3; ++c) std::swap(cur[c], prev[c]);
                    for (int c = 0; c < 3; ++c) *out++ = ((float)prev[c] * inv) * sc;
                } else {
                    for (int c = 0; c < 3; ++c) prev[c] = cur[c];
                }
                for (int c = 0; c < 3; ++c) *out++ = ((float)cur[c] * inv) * sc;
            }
        } else {
            for (int c = 0; c < 3; ++c) *out++ = ((float)cur[c] * inv) * sc;
        }
        smallidx += is_smaller;
        if (smallidx < kFirstIdx || smallidx >= kLastIdx) return fail(MB200_EINVAL, "corrupt XTC frame at byte %zu (smallidx %d)", offset, smallidx);
        if (is_smaller < 0) {
            smallnum = smaller;
            smaller = smallidx > kFirstIdx ? kMagicInts[smallidx - 1] / 2 : 0;
        } else if (is_smaller > 0) {
            smaller = smallnum;
            smallnum = kMagicInts[smallidx] / 2;
        }
        sizesmall[0] = sizesmall[1] = sizesmall[2] = (uint32_t)kMagicInts[smallidx];
        if (br.overrun) return fail(MB200_EINVAL, "corrupt XTC frame at byte %zu (bit stream too short)", offset);
    }
    return MB200_OK;
}
