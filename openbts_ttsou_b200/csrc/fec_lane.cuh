// fec_lane.cuh -- the transmit-side L1 encoders, ONE LANE PER BLOCK with every bit position known at compile time.
//
// fec_kernels.cuh's first encoder kernels spend a warp on a block and ~760 warp-instructions per frame moving single bytes
// (ncu: issue-active 86 %, 12 % of the HBM peak).  The work is integer and the positions never change, so here a lane keeps a
// whole block in registers as bit-packed words: the frame's 184 bytes are packed eight at a time with one multiply each, the
// Fire-code word comes from a byte-wise CRC table (shared memory), the convolutional code is two word-wide XORs of shifted
// copies (G0 = 1 + D^3 + D^4, G1 = 1 + D + D^3 + D^4: BitVector.cpp:290-338's generators 0x19 / 0x1b), and the 592 output bytes
// of a group of four bursts are assembled straight-line, each from ITS bit of the code planes -- interleaver, mapping on the
// burst and fixed fields folded into one compile-time table (EncTable) -- and stored as 16-byte words.  ~70 instructions per
// frame instead of ~760.  Same bits as fec.cuh's sequential forms (tests/test_fec_encode.py runs both against the reference).
#pragma once
#include <string.h>
#include <utility>

#include "fec.cuh"

namespace btsdsp {

// Output byte o = B * 148 + pos of a group of four bursts -> where its bit comes from: an index into the coded bits, or
// kEncSpecial | s for the fixed fields -- bit s of a per-group word: s = 0 a zero (tails), 1 a one, 2 Hl, 3 Hu, 4 + i midamble bit i.
constexpr unsigned short kEncSpecial = 0x8000;
struct alignas(16) EncTable { unsigned short idx[4 * 148]; };
__host__ __device__ constexpr unsigned short enc_fixed_field(int pos) {
  if (pos < 3 || pos >= 145) return kEncSpecial | 0;
  if (pos == 60) return kEncSpecial | 2;
  if (pos == 87) return kEncSpecial | 3;
  if (pos >= 61 && pos < 87) return (unsigned short)(kEncSpecial | (4 + pos - 61));
  return 0xffff;                                                         // an e-bit
}
// XCCH: c[k] goes to burst k % 4, e-bit 2*((49 k) % 57) + (k % 8)/4 (interleave, GSML1FEC.cpp:811-819)
__host__ __device__ constexpr EncTable make_xcch_table() {
  EncTable t{};
  for (int B = 0; B < 4; B++)
    for (int pos = 0; pos < 148; pos++) t.idx[B * 148 + pos] = enc_fixed_field(pos);
  for (int k = 0; k < 456; k++) {
    const int B = k % 4, j = 2 * ((49 * k) % 57) + ((k % 8) / 4);
    t.idx[B * 148 + (j < 57 ? 3 + j : 88 + (j - 57))] = (unsigned short)k;
  }
  return t;
}
// TCH: burst B of group g takes the even e-bits from block g (c[k], k % 8 == B: index 456 + k) and the odd ones from block g - 1
// (k % 8 == B + 4: index k)
__host__ __device__ constexpr EncTable make_tch_table() {
  EncTable t{};
  for (int B = 0; B < 4; B++)
    for (int pos = 0; pos < 148; pos++) t.idx[B * 148 + pos] = enc_fixed_field(pos);
  for (int k = 0; k < 456; k++) {
    const int r = k % 8, j = 2 * ((49 * k) % 57) + (r / 4);
    const int pos = j < 57 ? 3 + j : 88 + (j - 57);
    if (r < 4) t.idx[r * 148 + pos] = (unsigned short)(456 + k);
    else t.idx[(r - 4) * 148 + pos] = (unsigned short)k;
  }
  return t;
}
struct XcchTab { static constexpr EncTable tab = make_xcch_table(); };
struct TchTab { static constexpr EncTable tab = make_tch_table(); };

// byte-wise table of the 40-bit Fire-code encoder (Generator::encoderShift, BitVector.h:80-85, eight shifts at a time, first bit =
// the byte's MSB): state' = ((state << 8) & mask) ^ t[(state >> 32) ^ byte]
struct CrcTable { unsigned long long t[256]; };
constexpr unsigned long long kFireMask = (1ULL << 40) - 1;
__host__ __device__ constexpr CrcTable make_fire_crc_table() {
  CrcTable c{};
  for (int b = 0; b < 256; b++) {
    unsigned long long state = 0;
    for (int i = 7; i >= 0; i--) {
      const unsigned long long fb = ((state >> 39) ^ (unsigned long long)((b >> i) & 1)) & 1ULL;
      state = (state << 1) & kFireMask;
      if (fb) state ^= 0x10004820009ULL & kFireMask;
    }
    c.t[b] = state;
  }
  return c;
}

BTS_HD unsigned rev8(unsigned x) {
#ifdef __CUDA_ARCH__
  return __brev(x) >> 24;
#else
  x = ((x & 0xf0u) >> 4) | ((x & 0x0fu) << 4);
  x = ((x & 0xccu) >> 2) | ((x & 0x33u) << 2);
  return ((x & 0xaau) >> 1) | ((x & 0x55u) << 1);
#endif
}
BTS_HD unsigned rev32(unsigned x) {
#ifdef __CUDA_ARCH__
  return __brev(x);
#else
  return (rev8(x & 0xffu) << 24) | (rev8((x >> 8) & 0xffu) << 16) | (rev8((x >> 16) & 0xffu) << 8) | rev8(x >> 24);
#endif
}
BTS_HD unsigned popc32(unsigned x) {
#ifdef __CUDA_ARCH__
  return (unsigned)__popc(x);
#else
  return (unsigned)__builtin_popcount(x);
#endif
}
// four bytes (one bit each, value in bit 0) -> four bits, byte t at bit t: the products land on bits 24..27 and nowhere else
BTS_HD unsigned pack4(unsigned x) { return (((x & 0x01010101u) * 0x01020408u) >> 24) & 0xfu; }
BTS_HD unsigned load32(const unsigned char *p) {
#ifdef __CUDA_ARCH__
  return *reinterpret_cast<const unsigned *>(p);                         // global or shared memory (the staged kernel reads its tile)
#else
  unsigned v; memcpy(&v, p, 4); return v;
#endif
}

// u[0..228) of an XCCH / FACCH block as words (bit i of u at W[i >> 5], bit i & 31): the 184-bit frame (optionally LSB8MSB), the
// inverted Fire-code word MSB first, four tail zeros.  frame must be 4-byte aligned.
BTS_HD void xcch_u_words(const unsigned char *frame, int lsb8msb, const unsigned long long *crc, unsigned W[8]) {
#pragma unroll
  for (int j = 0; j < 8; j++) W[j] = 0;
  unsigned long long st = 0;
#pragma unroll
  for (int g = 0; g < 23; g++) {
    const unsigned n8 = pack4(load32(frame + 8 * g)) | (pack4(load32(frame + 8 * g + 4)) << 4);   // frame bit 8g + t at bit t
    const unsigned r8 = rev8(n8);
    const unsigned ub = lsb8msb ? r8 : n8;                               // u[8g + t] at bit t (LSB8MSB: u[8g+t] = frame[8g+7-t])
    const unsigned first_msb = lsb8msb ? n8 : r8;                        // the same eight u bits, u[8g] in bit 7: how the encoder eats them
    W[g >> 2] |= ub << (8 * (g & 3));
    const unsigned idx = ((unsigned)(st >> 32) ^ first_msb) & 0xffu;
    st = ((st << 8) & kFireMask) ^ crc[idx];
  }
  const unsigned long long p = ~st & kFireMask;                          // writeParityWord: inverted, MSB first at u[184..224)
  const unsigned q_lo = rev8((unsigned)(p >> 32) & 0xffu);               // u[184..192) = p bits 39..32
  const unsigned q_hi = rev32((unsigned)p);                              // u[192..224) = p bits 31..0
  W[5] |= q_lo << 24;
  W[6] = q_hi;
  W[7] = 0;
}
// The same without the table and its serial chain: the Fire code is linear, so parity bit j is the parity of the frame bits
// selected by mask j (mask bit i = bit j of the state a lone 1 at position i leaves behind): 40 x 6 independent AND + POPC.
// Measured in k_xcch_encode_tiles: 0.214 ms against 0.218 ms per 2^20 frames with the table -- the chain is not what limits that
// kernel; the table form is the one the kernels use, this one stays as its independent check (tests/test_fec_encode.py).
struct FireMasks { unsigned m[40][6]; };
__host__ __device__ constexpr FireMasks make_fire_masks() {
  FireMasks t{};
  for (int i = 0; i < 184; i++) {
    unsigned long long state = 0x10004820009ULL & kFireMask;
    for (int s = i + 1; s < 184; s++) {
      const unsigned long long fb = (state >> 39) & 1ULL;
      state = (state << 1) & kFireMask;
      if (fb) state ^= 0x10004820009ULL & kFireMask;
    }
    for (int j = 0; j < 40; j++) if ((state >> j) & 1ULL) t.m[j][i >> 5] |= 1u << (i & 31);
  }
  return t;
}
BTS_HD void xcch_u_words_popc(const unsigned char *frame, int lsb8msb, unsigned W[8]) {
#pragma unroll
  for (int j = 0; j < 8; j++) W[j] = 0;
#pragma unroll
  for (int g = 0; g < 23; g++) {
    const unsigned n8 = pack4(load32(frame + 8 * g)) | (pack4(load32(frame + 8 * g + 4)) << 4);
    W[g >> 2] |= (lsb8msb ? rev8(n8) : n8) << (8 * (g & 3));
  }
  constexpr FireMasks FM = make_fire_masks();
  unsigned lo = 0, hi = 0;                                               // the state's bits 0..31 / 32..39
#pragma unroll
  for (int j = 0; j < 40; j++) {
    unsigned c = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) c ^= popc32(W[i] & FM.m[j][i]);
    if (j < 32) lo |= (c & 1u) << j;
    else hi |= (c & 1u) << (j - 32);
  }
  W[5] |= rev8(~hi & 0xffu) << 24;                                       // inverted, MSB first at u[184..224)
  W[6] = rev32(~lo);
  W[7] = 0;
}
// code planes: bit k of G0 / G1 = c[2k] / c[2k+1] = u[k] ^ u[k-3] ^ u[k-4] / u[k] ^ u[k-1] ^ u[k-3] ^ u[k-4]
BTS_HD void conv_planes(const unsigned W[8], unsigned G0[8], unsigned G1[8]) {
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const unsigned lo = j ? W[j - 1] : 0u;
    const unsigned s1 = (W[j] << 1) | (lo >> 31), s3 = (W[j] << 3) | (lo >> 29), s4 = (W[j] << 4) | (lo >> 28);
    const unsigned a = W[j] ^ s3 ^ s4;
    G0[j] = a;
    G1[j] = a ^ s1;
  }
}

// pl: code planes as [block: 0 = table indices 0..455, 1 = 456..911][plane][8 words]; sp: the group's fixed-field word
template <unsigned short K>
BTS_HD unsigned enc_src_bit(const unsigned *pl, unsigned sp) {
  if constexpr ((K & kEncSpecial) != 0) {
    return (sp >> (K & 31)) & 1u;
  } else {
    constexpr int blk = K >= 456 ? 1 : 0, k = K - 456 * blk, pos = k >> 1;
    return (pl[(blk * 2 + (k & 1)) * 8 + (pos >> 5)] >> (pos & 31)) & 1u;
  }
}
template <class TAB, int W>
BTS_HD unsigned enc_out_word(const unsigned *pl, unsigned sp) {
  return enc_src_bit<TAB::tab.idx[4 * W]>(pl, sp) | (enc_src_bit<TAB::tab.idx[4 * W + 1]>(pl, sp) << 8) |
         (enc_src_bit<TAB::tab.idx[4 * W + 2]>(pl, sp) << 16) | (enc_src_bit<TAB::tab.idx[4 * W + 3]>(pl, sp) << 24);
}
template <class TAB, int Q>
BTS_HD void enc_out_quad(const unsigned *pl, unsigned sp, unsigned char *out) {
  const unsigned a = enc_out_word<TAB, 4 * Q>(pl, sp), b = enc_out_word<TAB, 4 * Q + 1>(pl, sp), c = enc_out_word<TAB, 4 * Q + 2>(pl, sp),
                 d = enc_out_word<TAB, 4 * Q + 3>(pl, sp);
#ifdef __CUDA_ARCH__
  *reinterpret_cast<uint4 *>(out + 16 * Q) = make_uint4(a, b, c, d);
#else
  const unsigned v[4] = {a, b, c, d};
  memcpy(out + 16 * Q, v, 16);
#endif
}
template <class TAB, int... Q>
BTS_HD void enc_out_all(const unsigned *pl, unsigned sp, unsigned char *out, std::integer_sequence<int, Q...>) {
  (enc_out_quad<TAB, Q>(pl, sp, out), ...);
}
// quads Q0 .. Q0 + N - 1 of a group (16 bytes each), written to out[0 .. 16 N): a piece of the group for a staging row
template <class TAB, int Q0, int... I>
BTS_HD void enc_out_range_(const unsigned *pl, unsigned sp, unsigned char *out, std::integer_sequence<int, I...>) {
  (enc_out_quad<TAB, Q0 + I>(pl, sp, out - 16 * Q0), ...);
}
template <class TAB, int Q0, int N>
BTS_HD void enc_out_range(const unsigned *pl, unsigned sp, unsigned char *out) {
  enc_out_range_<TAB, Q0>(pl, sp, out, std::make_integer_sequence<int, N>());
}
// the four bursts of a group (592 contiguous bytes: burst pitch 148), out 16-byte aligned
template <class TAB>
BTS_HD void enc_out_group(const unsigned *pl, unsigned sp, unsigned char *out) {
  enc_out_all<TAB>(pl, sp, out, std::make_integer_sequence<int, 37>());
}

// ---- traffic channel ----
// 3-bit class-1A parity (Parity(0x0b, 3, 50)) as three masks: state bit j = parity of the d bits selected by mask j (the code is
// linear; mask bit i = bit j of the state a lone 1 at position i leaves behind)
struct TchParMasks { unsigned m[3][2]; };
__host__ __device__ constexpr TchParMasks make_tch_par_masks() {
  TchParMasks t{};
  for (int i = 0; i < 50; i++) {
    unsigned state = 0x0bu & 7u;
    for (int s = i + 1; s < 50; s++) {
      const unsigned fb = (state >> 2) & 1u;
      state = (state << 1) & 7u;
      if (fb) state ^= 0x0bu & 7u;
    }
    for (int j = 0; j < 3; j++) if ((state >> j) & 1u) t.m[j][i >> 5] |= 1u << (i & 31);
  }
  return t;
}
// the even-position bits of x, packed into the low 16 bits
BTS_HD unsigned even16(unsigned x) {
  x &= 0x55555555u;
  x = (x | (x >> 1)) & 0x33333333u;
  x = (x | (x >> 2)) & 0x0f0f0f0fu;
  x = (x | (x >> 4)) & 0x00ff00ffu;
  return (x | (x >> 8)) & 0xffffu;
}
// code planes of one speech frame d[260] (class order; encodeTCH, GSML1FEC.cpp:1248-1279): u[k] = d[2k], u[184-k] = d[2k+1]
// (k <= 90), u[91..94) = the inverted 3-bit parity of d[0..50) MSB first, four tail zeros, class 1 coded; the class-2 bits
// d[182..260) follow uncoded as c[378..456).  All of it on words: e / o = the even / odd bits of d.  d must be 4-byte aligned.
BTS_HD void tch_speech_planes(const unsigned char *d, unsigned G0[8], unsigned G1[8]) {
  unsigned D[10];
#pragma unroll
  for (int j = 0; j < 10; j++) D[j] = 0;
#pragma unroll
  for (int i = 0; i < 65; i++) D[i >> 3] |= pack4(load32(d + 4 * i)) << (4 * (i & 7));
  unsigned E[5], O[5];
#pragma unroll
  for (int j = 0; j < 5; j++) {
    E[j] = even16(D[2 * j]) | (even16(D[2 * j + 1]) << 16);
    O[j] = even16(D[2 * j] >> 1) | (even16(D[2 * j + 1] >> 1) << 16);
  }
  constexpr TchParMasks PM = make_tch_par_masks();
  unsigned state = 0;
#pragma unroll
  for (int j = 0; j < 3; j++) state |= ((popc32(D[0] & PM.m[j][0]) ^ popc32(D[1] & PM.m[j][1])) & 1u) << j;
  const unsigned p = ~state & 7u;
  constexpr unsigned LOW27 = (1u << 27) - 1;
  const unsigned R0 = rev32(O[2] & LOW27), R1 = rev32(O[1]), R2 = rev32(O[0]);      // o[0..91) reversed over 96 bits
  unsigned W[8];
  W[0] = E[0];
  W[1] = E[1];
  W[2] = (E[2] & LOW27) | (((p >> 2) & 1u) << 27) | (((p >> 1) & 1u) << 28) | ((p & 1u) << 29) | (R0 << 25);
  W[3] = (R1 << 25) | (R0 >> 7);
  W[4] = (R2 << 25) | (R1 >> 7);
  W[5] = R2 >> 7;
  W[6] = 0;
  W[7] = 0;
  conv_planes(W, G0, G1);
  // class 2: plane bit 189 + m = e[91 + m] / o[91 + m], m < 39: the bits from 91 up, moved up by 98
  const unsigned e2 = E[2] & ~LOW27, o2 = O[2] & ~LOW27;
  G0[5] |= e2 << 2;
  G0[6] |= (E[3] << 2) | (e2 >> 30);
  G0[7] |= (E[4] << 2) | (E[3] >> 30);
  G1[5] |= o2 << 2;
  G1[6] |= (O[3] << 2) | (o2 >> 30);
  G1[7] |= (O[4] << 2) | (O[3] >> 30);
}
// code planes of block q of a traffic channel: a stolen block is a FACCH frame coded like an XCCH block (dispatch :1323-1333)
BTS_HD void tch_block_planes(int stolen, const unsigned char *d, const unsigned char *f, int lsb8msb, const unsigned long long *crc,
                             unsigned G0[8], unsigned G1[8]) {
  if (stolen) {
    unsigned W[8];
    xcch_u_words(f, lsb8msb, crc, W);
    conv_planes(W, G0, G1);
  } else {
    tch_speech_planes(d, G0, G1);
  }
}
// group g of four bursts from the planes of block g - 1 (pl[0..16): odd e-bits, Hl) and block g (pl[16..32): even e-bits, Hu)
BTS_HD void tch_encode_group_lane(const unsigned *pl, int prev_stolen, int cur_stolen, unsigned sp_base, unsigned char *bursts) {
  enc_out_group<TchTab>(pl, sp_base | ((unsigned)(prev_stolen != 0) << 2) | ((unsigned)(cur_stolen != 0) << 3), bursts);
}

// one XCCH frame, lane form
BTS_HD void xcch_encode_frame_lane(const unsigned char *frame, int lsb8msb, const unsigned long long *crc, unsigned sp_base,
                                   unsigned char *bursts) {
  unsigned W[8], pl[16];
  xcch_u_words(frame, lsb8msb, crc, W);
  conv_planes(W, pl, pl + 8);
  enc_out_group<XcchTab>(pl, sp_base | (1u << 2) | (1u << 3), bursts);       // Hl = Hu = 1, GSML1FEC.cpp:735-736
}

}  // namespace btsdsp
