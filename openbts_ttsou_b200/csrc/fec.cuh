// fec.cuh -- L1 FEC after the receive path: the XCCH block decoder of GSM 05.03 4.1 as OpenBTS runs
// it: XCCHL1Decoder::deinterleave + decode (GSML1FEC.cpp:616-660) = deinterleave 4 x 114 e-bits -> c[456], soft-input
// Viterbi for the rate-1/2, K = 5 code (SoftVector::decode + ViterbiR2O4, CommonLibs/BitVector.cpp:290-540) -> u[228],
// 40-bit Fire-code syndrome over d[184]:p[40] with the parity inverted (Parity/Generator, BitVector.h:39-112).
// Input = the RX datagrams' soft bytes; the socket side turns them into probabilities as byte / 256.0F
// (ARFCNManager::driveRx, TRXManager.cpp:230).
//
// Parallel shape: one WARP per frame, one LANE per trellis candidate -- the reference's decoder keeps 16 survivors and
// 32 candidates per step, so branch / metric / prune are one instruction each across the warp, and the path costs are
// accumulated per lane with exactly the reference's two float additions per step (ties and the first-minimum rule
// then resolve identically).  xcch_decode_frame_seq is the same algorithm as a plain loop (host emulation, tests).
#pragma once
#include <stdint.h>

#include "cplx.cuh"

namespace btsdsp {

constexpr int kVitStates = 16, kVitCands = 32, kVitDeferral = 24;        // BitVector.h:128-138 (order 4, deferral 6*order)
constexpr int kXcchC = 456, kXcchU = 228, kXcchSteps = kXcchU + kVitDeferral, kXcchTable = 2 * kXcchSteps;   // 504

// generator output for a 5-bit input history: (g0 << 1) | g1 with g0 = 0x19, g1 = 0x1b (BitVector.cpp:290-338)
BTS_HD unsigned vit_parity5(unsigned v) { v ^= v >> 4; v ^= v >> 2; v ^= v >> 1; return v & 1u; }
BTS_HD unsigned vit_generator(unsigned hist5) { return (vit_parity5(hist5 & 0x19u) << 1) | vit_parity5(hist5 & 0x1bu); }
// the same as a 64-bit lookup word: entry h in bits [2h, 2h+2)
__host__ __device__ constexpr unsigned long long vit_generator_lut() {
  unsigned long long w = 0;
  for (unsigned h = 0; h < 32; h++) {
    unsigned a = h & 0x19u, b = h & 0x1bu;
    a ^= a >> 4; a ^= a >> 2; a ^= a >> 1;
    b ^= b >> 4; b ^= b >> 2; b ^= b >> 1;
    w |= (unsigned long long)(((a & 1u) << 1) | (b & 1u)) << (2 * h);
  }
  return w;
}

// GSM 05.03 4.1.4 as the reference writes it (GSML1FEC.cpp:620-624): c[k] = i[k % 4][2*((49 k) % 57) + ((k % 8) / 4)],
// and i[B][j] = burst B's bit 3 + j (j < 57) or 88 + (j - 57) (:603-604)
BTS_HD int xcch_source_bit(int k, int *B) {
  *B = k & 3;
  const int j = 2 * ((49 * k) % 57) + ((k & 7) >> 2);
  return j < 57 ? 3 + j : 88 + (j - 57);
}

// the per-bit tables of SoftVector::decode (BitVector.cpp:476-497) for probability p = byte / 256
BTS_HD void vit_costs(float p, float *match, float *mismatch, unsigned *hard) {
  *hard = p > 0.5F ? 1u : 0u;                                            // sliced(), :440-448
  float pVal = p;
  if (pVal > 0.5F) pVal = BTS_SUB(1.0F, pVal);
  float ipVal = BTS_SUB(1.0F, pVal);
  if (pVal < 0.01F) pVal = (float)0.01;
  if (ipVal < 0.01F) ipVal = (float)0.01;
  *match = BTS_DIV(0.25F, ipVal);
  *mismatch = BTS_DIV(0.25F, pVal);
}

// Fire-code syndrome (Generator::syndromeShift, BitVector.h:69-74) over d[184] : inverted p[40]
BTS_HD bool xcch_parity_ok(const unsigned char *u) {
  const unsigned long long coeff = 0x10004820009ULL, mask = (1ULL << 40) - 1;
  unsigned long long state = 0;
  for (int i = 0; i < 224; i++) {
    const unsigned bit = (i < 184 ? u[i] : ~u[i]) & 1u;                  // mP.invert(), GSML1FEC.cpp:649
    const unsigned fb = (unsigned)(state >> 39) & 1u;
    state = (state << 1) ^ bit;
    if (fb) state ^= coeff;
  }
  // XCCHL1Decoder::decode keeps the 64-bit syndrome in an `unsigned` (GSML1FEC.cpp:652): only its low 32 bits are tested
  return (unsigned)(state & mask) == 0u;
}

// SoftVector::decode (BitVector.cpp:451-540), sequential form: NC coded probabilities -> NU = NC/2 decoded bits
template <int NC, int NU>
BTS_HD void viterbi_seq(const float *p, unsigned char *u) {
  constexpr int STEPS = NU + kVitDeferral, TABLE = 2 * STEPS;
  float match[TABLE], mismatch[TABLE];
  unsigned char hard[TABLE];
  for (int k = 0; k < NC; k++) {
    unsigned h;
    vit_costs(p[k], &match[k], &mismatch[k], &h);
    hard[k] = (unsigned char)h;
  }
  for (int k = NC; k < TABLE; k++) { match[k] = 0.5F; mismatch[k] = 0.5F; hard[k] = hard[NC - 1]; }   // :462-466, :492-496
  float scost[kVitStates];
  unsigned sin_[kVitStates], sout[kVitStates];
  for (int i = 0; i < kVitStates; i++) { scost[i] = 0.0F; sin_[i] = 0; sout[i] = 0; }
  for (int s = 0; s < STEPS; s++) {
    float ccost[kVitCands];
    unsigned cin[kVitCands], cout[kVitCands];
    const unsigned in2 = ((unsigned)hard[2 * s] << 1) | hard[2 * s + 1];   // history[2s+1], low two bits
    for (int c = 0; c < kVitCands; c++) {                                // branchCandidates + getSoftCostMetrics
      const int sp = c >> 1;
      cin[c] = (sin_[sp] << 1) | (unsigned)(c & 1);
      cout[c] = (sout[sp] << 2) | vit_generator(cin[c] & 0x1fu);
      const unsigned mm = in2 ^ cout[c];
      const float t = BTS_ADD((mm & 1u) ? mismatch[2 * s + 1] : match[2 * s + 1], ((mm >> 1) & 1u) ? mismatch[2 * s] : match[2 * s]);
      ccost[c] = BTS_ADD(scost[sp], t);
    }
    for (int i = 0; i < kVitStates; i++) {                               // pruneCandidates
      const int w = ccost[i] < ccost[i + kVitStates] ? i : i + kVitStates;
      scost[i] = ccost[w]; sin_[i] = cin[w]; sout[i] = cout[w];
    }
    int best = 0;                                                        // minCost: first strict minimum
    for (int i = 1; i < kVitStates; i++) if (scost[i] < scost[best]) best = i;
    if (s >= kVitDeferral) u[s - kVitDeferral] = (unsigned char)((sin_[best] >> kVitDeferral) & 1u);
  }
}

// One XCCH frame.  soft: the four bursts' soft bytes (burst_pitch apart).  u: 228 bits out.
BTS_HD bool xcch_decode_frame_seq(const unsigned char *soft, int burst_pitch, unsigned char *u) {
  float p[kXcchC];
  for (int k = 0; k < kXcchC; k++) {
    int B;
    const int bit = xcch_source_bit(k, &B);
    p[k] = (float)soft[B * burst_pitch + bit] / 256.0F;
  }
  viterbi_seq<kXcchC, kXcchU>(p, u);
  return xcch_parity_ok(u);
}

// ---- TCH/FACCH (GSM 05.03 3.1, 4.2), TCHFACCHL1Decoder::processBurst / deinterleave / decodeTCH, GSML1FEC.cpp:1031-1210 ----
// A traffic channel's bursts are diagonally interleaved over EIGHT bursts: block q takes the even-position e-bits of the
// four bursts that complete it and the odd-position e-bits of the four before (deinterleave(blockOffset), :1102-1110:
// c[k] = i[(k + off) % 8][2*((49 k) % 57) + ((k % 8)/4)]; with the reference's B numbering this is burst 4q + (k % 8) of a
// stream in which block q occupies bursts 4q .. 4q+7).  The burst that completes the block carries the stealing flag Hl
// (bit 60, :1073): set -> the 456 bits are a FACCH frame, decoded exactly like an XCCH block (:1075-1084); clear -> a speech
// frame (:1129-1160): class 1 = c[0..378) through the same Viterbi -> u[189], class 2 = c[378..456) sliced; d[2k] = u[k],
// d[2k+1] = u[184-k] (k <= 90), d[182..260) = class 2; good = 3-bit parity of d[0..50) (generator 0x0b, sent inverted in
// u[91..94)) matches AND the four tail bits u[185..189) are zero.  The GSM 06.10 frame formatting and the bad-frame
// substitution that follow (:1161-1190, the latter draws from random()) belong to the speech layer and are not part of this.
constexpr int kTchC1 = 378, kTchU = 189, kTchC2 = 78, kTchD = 260;
BTS_HD int tch_source_bit(int k, int *burst) {
  const int r = k & 7;
  *burst = r;
  const int j = 2 * ((49 * k) % 57) + (r >> 2);
  return j < 57 ? 3 + j : 88 + (j - 57);
}
BTS_HD bool tch_fields(const unsigned char *u, const unsigned char *c2, unsigned char *d) {
  for (int k = 0; k <= 90; k++) { d[2 * k] = u[k] & 1; d[2 * k + 1] = u[184 - k] & 1; }   // :1141-1144
  for (int i = 0; i < kTchC2; i++) d[182 + i] = c2[i] & 1;                                 // :1137
  const unsigned sent = (~(((unsigned)(u[91] & 1) << 2) | ((unsigned)(u[92] & 1) << 1) | (unsigned)(u[93] & 1))) & 7u;   // :1148
  unsigned state = 0;                                                                      // mClass1A_d.parity(Parity(0x0b,3,50))
  for (int i = 0; i < 50; i++) {
    const unsigned fb = ((state >> 2) ^ d[i]) & 1u;
    state <<= 1;
    if (fb) state ^= 0x0bu;
  }
  unsigned tail = 0;
  for (int i = 0; i < 4; i++) tail = (tail << 1) | (u[185 + i] & 1u);                      // :1153
  return sent == (state & 7u) && tail == 0;                                                // :1160
}
// One block, sequential form (host emulation, tests).  soft: the block's eight bursts.  Outputs: d[260] + *good when the
// frame is not stolen; fu[228] + *fok when it is (the other pair is zeroed).  Returns the stolen flag.
BTS_HD bool tch_decode_block_seq(const unsigned char *soft, int burst_pitch, unsigned char *d, int *good, unsigned char *fu, int *fok) {
  float p[kXcchC];
  for (int k = 0; k < kXcchC; k++) {
    int B;
    const int bit = tch_source_bit(k, &B);
    p[k] = (float)soft[B * burst_pitch + bit] / 256.0F;
  }
  const bool stolen = (float)soft[7 * burst_pitch + 60] / 256.0F > 0.5F;                   // inBurst.Hl(), bit(gHlIndex)
  for (int i = 0; i < kTchD; i++) d[i] = 0;
  for (int i = 0; i < kXcchU; i++) fu[i] = 0;
  *good = 0; *fok = 0;
  if (stolen) {
    viterbi_seq<kXcchC, kXcchU>(p, fu);
    *fok = xcch_parity_ok(fu) ? 1 : 0;
  } else {
    unsigned char u[kTchU], c2[kTchC2];
    viterbi_seq<kTchC1, kTchU>(p, u);
    for (int i = 0; i < kTchC2; i++) c2[i] = p[kTchC1 + i] > 0.5F ? 1 : 0;                 // sliced()
    *good = tch_fields(u, c2, d) ? 1 : 0;
  }
  return stolen;
}

// ---- RACH (GSM 05.03 4.6), RACHL1Decoder::writeLowSide, GSML1FEC.cpp:474-515 ----
// e = burst bits 49..84 (36 coded bits) -> u[18] = d[8] : p[6] : tail[4].  The caller checks tail == 0 and
// bsic == its BSIC (the parity word is sent inverted and XORed with the BSIC); ra = the 8-bit RA field.
constexpr int kRachC = 36, kRachU = 18;
BTS_HD void rach_fields(const unsigned char *u, int *tail, int *bsic, int *ra) {
  int t = 0, sent = 0, r = 0;
  for (int i = 0; i < 4; i++) t = (t << 1) | (u[14 + i] & 1);            // peekField(14,4) :485
  for (int i = 0; i < 6; i++) sent = (sent << 1) | (u[8 + i] & 1);       // peekField(8,6)
  unsigned state = 0;                                                    // mD.parity(Parity(0x06f,6,8)): encoderShift x 8
  for (int i = 0; i < 8; i++) {
    const unsigned fb = ((state >> 5) ^ u[i]) & 1u;
    state <<= 1;
    if (fb) state ^= 0x06fu;
  }
  *tail = t;
  *bsic = (int)(((unsigned)~sent ^ (state & 0x3fu)) & 0x3fu);            // :492-494
  for (int i = 0; i < 8; i++) r |= (u[i] & 1) << i;                      // LSB8MSB then peekField(0,8) :507-508
  *ra = r;
}
BTS_HD void rach_decode_burst_seq(const unsigned char *soft, unsigned char *u, int *tail, int *bsic, int *ra) {
  float p[kRachC];
  for (int k = 0; k < kRachC; k++) p[k] = (float)soft[49 + k] / 256.0F;
  viterbi_seq<kRachC, kRachU>(p, u);
  rach_fields(u, tail, bsic, ra);
}

// ---- L1 encoders on the transmit side: what produces the 148-bit bursts modulateBurst is called with -----------------
// XCCHL1Encoder::sendFrame / encode / interleave / transmit (GSML1FEC.cpp:763-850) and TCHFACCHL1Encoder::encodeTCH / dispatch /
// interleave (:1248-1392).  Building blocks as the reference's classes compute them: BitVector::LSB8MSB (BitVector.cpp:189-195,
// every whole byte bit-reversed), Generator::encoderShift + Parity::writeParityWord (BitVector.h:80-85, BitVector.cpp:411-416: the
// remainder, inverted, written MSB first), BitVector::encode (BitVector.cpp:217-238: c[2i], c[2i+1] = the two generator outputs for
// the five-bit history ending at u[i]), the block (XCCH) or diagonal (TCH) interleaver, and the normal burst's fixed fields
// (GSMTransfer.h:47-48: Hl = bit 60, Hu = bit 87; midamble at 61..86; tails and guard zero).
BTS_HD int lsb8msb_src(int i) { return (i & ~7) + 7 - (i & 7); }          // source index of bit i after LSB8MSB (whole bytes only)
BTS_HD unsigned long long fec_parity_seq(const unsigned char *d, int n, unsigned long long coeff, int len) {
  unsigned long long state = 0;
  for (int i = 0; i < n; i++) {
    const unsigned fb = ((unsigned)(state >> (len - 1)) ^ d[i]) & 1u;
    state <<= 1;
    if (fb) state ^= coeff;
  }
  return state;
}
BTS_HD void conv_encode_seq(const unsigned char *u, int n, unsigned char *c) {
  unsigned acc = 0;
  for (int i = 0; i < n; i++) {
    acc = (acc << 1) | (u[i] & 1u);
    const unsigned g = vit_generator(acc & 0x1fu);
    c[2 * i] = (unsigned char)(g >> 1);
    c[2 * i + 1] = (unsigned char)(g & 1u);
  }
}
// tsc_word: midamble bit i (burst bit 61 + i) in bit 25 - i; have_tsc == 0 leaves 61..86 zero for the caller to fill
BTS_HD unsigned char burst_tsc_bit(unsigned tsc_word, int have_tsc, int pos) {
  return (unsigned char)(have_tsc ? (tsc_word >> (25 - (pos - 61))) & 1u : 0u);
}
// the fixed-field word of the table-driven kernels (fec_kernels.cuh): bit 1 = a one, bits 4 + i = midamble bit i (Hl, Hu: bits 2, 3)
BTS_HD unsigned enc_sp_base(unsigned tsc_word, int have_tsc) {
  unsigned sp = 1u << 1;
  if (have_tsc) for (int i = 0; i < 26; i++) sp |= ((tsc_word >> (25 - i)) & 1u) << (4 + i);
  return sp;
}
// u[228] of an XCCH / FACCH block from its 184-bit L2 frame: d (optionally LSB8MSB), inverted Fire-code parity, four tail zeros
BTS_HD void xcch_build_u_seq(const unsigned char *frame, int lsb8msb, unsigned char *u) {
  for (int i = 0; i < 184; i++) u[i] = frame[lsb8msb ? lsb8msb_src(i) : i] & 1;
  const unsigned long long p = ~fec_parity_seq(u, 184, 0x10004820009ULL, 40);
  for (int j = 0; j < 40; j++) u[184 + j] = (unsigned char)((p >> (39 - j)) & 1ULL);
  for (int j = 224; j < 228; j++) u[j] = 0;
}
// One XCCH frame -> four normal bursts (148 bits each, burst_pitch apart), stealing flags set (:735-738), sequential form
BTS_HD void xcch_encode_frame_seq(const unsigned char *frame, int lsb8msb, unsigned tsc_word, int have_tsc, unsigned char *bursts,
                                  int burst_pitch) {
  unsigned char u[kXcchU], c[kXcchC];
  xcch_build_u_seq(frame, lsb8msb, u);
  conv_encode_seq(u, kXcchU, c);
  for (int B = 0; B < 4; B++) {
    unsigned char *b = bursts + B * burst_pitch;
    for (int i = 0; i < 148; i++) b[i] = (i >= 61 && i < 87) ? burst_tsc_bit(tsc_word, have_tsc, i) : 0;
    b[60] = 1; b[87] = 1;
  }
  for (int k = 0; k < kXcchC; k++) {
    int B;
    const int pos = xcch_source_bit(k, &B);
    bursts[B * burst_pitch + pos] = c[k];
  }
}
// c[456] of one traffic-channel block: a speech frame d[260] (class order, encodeTCH :1248-1279) or a FACCH frame f[184] (:1323-1333)
BTS_HD void tch_encode_c_seq(int stolen, const unsigned char *d, const unsigned char *f, int lsb8msb, unsigned char *c) {
  if (stolen) {
    unsigned char u[kXcchU];
    xcch_build_u_seq(f, lsb8msb, u);
    conv_encode_seq(u, kXcchU, c);
    return;
  }
  unsigned char u[kTchU];
  for (int k = 0; k <= 90; k++) { u[k] = d[2 * k] & 1; u[184 - k] = d[2 * k + 1] & 1; }
  const unsigned p = ~(unsigned)fec_parity_seq(d, 50, 0x0bULL, 3);            // d holds bits (0/1) -- masked below
  for (int j = 0; j < 3; j++) u[91 + j] = (unsigned char)((p >> (2 - j)) & 1u);
  for (int k = 185; k <= 188; k++) u[k] = 0;
  conv_encode_seq(u, kTchU, c);
  for (int i = 0; i < kTchC2; i++) c[kTchC1 + i] = d[182 + i] & 1;
}
// A traffic channel's block stream, sequential form: nblocks blocks -> 4*nblocks + 4 bursts.  Block q fills the even e-bits and Hu
// of bursts 4q..4q+3 and the odd e-bits and Hl of bursts 4q+4..4q+7 (dispatch :1355-1381 with mOffset alternating, Hu = this
// block's stealing flag, Hl = the previous block's).  carry = the four half-filled closing bursts of the previous call (their odd
// e-bits and Hl are taken over) or NULL for a channel that starts here (the constructor's zero-filled interleaver, :1219-1224).
BTS_HD void tch_encode_stream_seq(const unsigned char *d260, const unsigned char *f184, const unsigned char *steal, long long nblocks,
                                  int lsb8msb, unsigned tsc_word, int have_tsc, const unsigned char *carry, unsigned char *bursts,
                                  int burst_pitch) {
  for (long long b = 0; b < 4 * nblocks + 4; b++) {
    unsigned char *bp = bursts + b * burst_pitch;
    for (int i = 0; i < 148; i++) bp[i] = (i >= 61 && i < 87) ? burst_tsc_bit(tsc_word, have_tsc, i) : 0;
  }
  if (carry)
    for (int B = 0; B < 4; B++) {
      bursts[B * burst_pitch + 60] = carry[B * burst_pitch + 60] & 1;
      for (int k = 0; k < kXcchC; k++) {
        int r;
        const int pos = tch_source_bit(k, &r);
        if (r == B + 4) bursts[B * burst_pitch + pos] = carry[B * burst_pitch + pos] & 1;
      }
    }
  for (long long q = 0; q < nblocks; q++) {
    unsigned char c[kXcchC];
    const int st = steal[q] ? 1 : 0;
    tch_encode_c_seq(st, d260 + q * kTchD, f184 + q * 184, lsb8msb, c);
    for (int k = 0; k < kXcchC; k++) {
      int r;
      const int pos = tch_source_bit(k, &r);
      bursts[(4 * q + r) * burst_pitch + pos] = c[k];
    }
    for (int B = 0; B < 4; B++) {
      bursts[(4 * q + B) * burst_pitch + 87] = (unsigned char)st;
      bursts[(4 * q + 4 + B) * burst_pitch + 60] = (unsigned char)st;
    }
  }
}

}  // namespace btsdsp
