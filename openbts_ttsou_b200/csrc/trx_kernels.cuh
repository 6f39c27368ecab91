// trx_kernels.cuh -- kernels of the caller-policy pipeline (see trx_policy.cuh); included by kernels.cu.
// A batch is nframes x narfcn x 8 bursts laid out [frame][arfcn][tn] at a fixed pitch; FN = fn0 + frame.

struct __align__(16) DfeRec {           // a designed equaliser, by the index of the burst it was estimated from
  cf w[7];
  cf b[5];
  float off, pad[3];
};

// pass 2: one thread per ARFCN walks its bursts in FIFO order
__global__ void k_trx_policy(const DevTables *__restrict__ T, TrxState *__restrict__ st, int narfcn, int nframes, int fn0,
                             const DetRec *__restrict__ det, const int *__restrict__ rach_slot,
                             const int *__restrict__ rach_flag, int *__restrict__ act, double *__restrict__ thr_at,
                             int *__restrict__ commit, int need_dfe) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= narfcn) return;
  TrxScalars s;
  trx_load_scalars(st[a], s);
  int cm[8];
  trx_policy_arfcn(s, nframes, fn0, narfcn, a, det, rach_slot, rach_flag, T->exp_neg, act, thr_at, cm, need_dfe != 0);
  // w, b and chan_off of the state are committed by k_trx_commit once pass 3 has designed them
  trx_store_scalars(st[a], s);
#pragma unroll
  for (int tn = 0; tn < 8; tn++) commit[a * 8 + tn] = cm[tn];
}

// pass 3a: designDFE for the bursts that re-estimate (Transceiver.cpp:346-347)
__global__ void k_trx_design(long long n, const DetRec *__restrict__ det, const int *__restrict__ act,
                             const double *__restrict__ thr_at, DfeRec *__restrict__ dfe) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || act[i] != (int)i) return;
  const DetRec d = det[i];
  const cf ia = cdiv(mk(1.0F, 0.0F), mk(d.amp_x, d.amp_y));
  cf ch[6], w[7], fb[5];
#pragma unroll
  for (int j = 0; j < 6; j++) ch[j] = cmul(d.chan[j], ia);                               // scaleVector :346
  design_dfe<7, 5>(ch, 5, trx_snr_estimate(mk(d.amp_x, d.amp_y), thr_at[i]), 7, w, fb);   // :340, :347
  DfeRec r;
#pragma unroll
  for (int j = 0; j < 7; j++) r.w[j] = w[j];
#pragma unroll
  for (int j = 0; j < 5; j++) r.b[j] = fb[j];
  r.off = d.off;
  r.pad[0] = r.pad[1] = r.pad[2] = 0.0F;
  dfe[i] = r;
}

// pass 3b: the equaliser's per-burst parameters: 1/amplitude, TOA - chanRespOffset, the (possibly cached) taps (:391-396)
__global__ void k_trx_eqparams(long long n, int narfcn, const DetRec *__restrict__ det, const int *__restrict__ act,
                               const DfeRec *__restrict__ dfe, const TrxState *__restrict__ st, EqParams *__restrict__ eqp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  EqParams e;
  const int a = act[i];
  if (a >= 0 || a == ACT_CARRIED) {
    const DetRec d = det[i];
    const cf ia = cdiv(mk(1.0F, 0.0F), mk(d.amp_x, d.amp_y));
    float off;
    if (a >= 0) {
      const DfeRec r = dfe[a];
#pragma unroll
      for (int j = 0; j < 7; j++) e.w[j] = r.w[j];
#pragma unroll
      for (int j = 0; j < 5; j++) e.b[j] = r.b[j];
      off = r.off;
    } else {
      const int arfcn = (int)((i >> 3) % narfcn), tn = (int)(i & 7);
#pragma unroll
      for (int j = 0; j < 7; j++) e.w[j] = st[arfcn].w[tn][j];
#pragma unroll
      for (int j = 0; j < 5; j++) e.b[j] = st[arfcn].b[tn][j];
      off = st[arfcn].chan_off[tn];
    }
    e.ia_x = ia.x; e.ia_y = ia.y; e.toa_eq = BTS_SUB(d.toa, off); e.ok = 1.0F;
  } else {
    e.ia_x = e.ia_y = e.toa_eq = e.ok = 0.0F;
#pragma unroll
    for (int j = 0; j < 7; j++) e.w[j] = mk(0.0F, 0.0F);
#pragma unroll
    for (int j = 0; j < 5; j++) e.b[j] = mk(0.0F, 0.0F);
  }
  eqp[i] = e;
}

// pass 3b without the equaliser (second variant, need_dfe == false): every accepted burst -- normal or access -- goes through
// demodulateBurst with its own amplitude and TOA (Transceiver52M/Transceiver.cpp:382-386); one record per burst of the batch
__global__ void k_trx_slice_params(long long n, const DetRec *__restrict__ det, const int *__restrict__ act,
                                   const int *__restrict__ rach_slot, const cf *__restrict__ rach_amp,
                                   const float *__restrict__ rach_toa, EqParams *__restrict__ eqp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int a = act[i];
  float4 q = make_float4(0.0F, 0.0F, 0.0F, 0.0F);
  if (a == ACT_SLICE || a == ACT_RACH) {
    cf amp;
    float toa;
    if (a == ACT_RACH) { const int j = rach_slot[i]; amp = rach_amp[j]; toa = rach_toa[j]; }
    else { amp = mk(det[i].amp_x, det[i].amp_y); toa = det[i].toa; }
    const cf ia = cdiv(mk(1.0F, 0.0F), amp);                                             // ((complex) 1.0)/channel
    q = make_float4(ia.x, ia.y, toa, 1.0F);
  }
  reinterpret_cast<float4 *>(eqp + i)[0] = q;
}

// RACH bursts the policy did not accept (energy gate) must not be demodulated
__global__ void k_trx_rach_veto(long long nr, const int *__restrict__ rach_idx, const int *__restrict__ act,
                                EqParams *__restrict__ eqp_r) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nr) return;
  if (act[rach_idx[j]] != ACT_RACH) reinterpret_cast<float4 *>(eqp_r + j)[0].w = 0.0F;
}

// pass 3c: RX datagrams (Transceiver.cpp:400-402, 659-673): one thread per burst writes the 8 header bytes as two
// 4-byte stores.  The equaliser has already written the 148 soft bytes of the normal bursts
// (and zeros elsewhere) at dgram + 8; the few RACH bursts convert their soft bits from the compact float rows here.
__global__ void k_trx_datagram(const DevTables *__restrict__ T, long long n, int narfcn, int fn0, const DetRec *__restrict__ det, const int *__restrict__ act,
                               const int *__restrict__ rach_slot, const cf *__restrict__ rach_amp,
                               const float *__restrict__ rach_toa, const float *__restrict__ rach_soft, int rach_soft_pitch,
                               int *__restrict__ valid, unsigned char *__restrict__ dgram, int dgram_pitch,
                               const float *__restrict__ soft_by_burst = nullptr) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned char *dg = dgram + i * (long long)dgram_pitch;
  const int a = act[i];
  unsigned char hdr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (a != ACT_NONE) {
    const int tn = (int)(i & 7), fn = (int)((fn0 + i / (8LL * narfcn)) % kHyperframe);
    cf amp;
    float toa;
    if (a == ACT_RACH) { const int j = rach_slot[i]; amp = rach_amp[j]; toa = rach_toa[j]; }
    else { amp = mk(det[i].amp_x, det[i].amp_y); toa = det[i].toa; }
    if (a == ACT_RACH || soft_by_burst) {
      // soft floats -> datagram bytes: the compact access-burst rows, or (no-equaliser mode) this burst's own row
      const float *sp = soft_by_burst ? soft_by_burst + i * rach_soft_pitch : rach_soft + (long long)rach_slot[i] * rach_soft_pitch;
      for (int m = 0; m < 148; m += 4) {
        const unsigned w = trx_soft_byte(sp[m]) | (trx_soft_byte(sp[m + 1]) << 8) | (trx_soft_byte(sp[m + 2]) << 16) |
                           ((unsigned)trx_soft_byte(sp[m + 3]) << 24);
        *reinterpret_cast<unsigned *>(dg + 8 + m) = w;
      }
      if (soft_by_burst) *reinterpret_cast<unsigned *>(dg + 156) = 0u;                  // bytes 156..159 (the equaliser's job otherwise)
    }
    trx_datagram_header(T, hdr, tn, fn, amp, toa, 1);
  } else if (soft_by_burst) {
    for (int m = 8; m < 160; m += 4) *reinterpret_cast<unsigned *>(dg + m) = 0u;       // rows of invalid bursts are zero
  }
  unsigned h0 = 0, h1 = 0;                                   // rows are 4-byte aligned (dgram_pitch % 4 == 0)
#pragma unroll
  for (int k = 0; k < 4; k++) { h0 |= (unsigned)hdr[k] << (8 * k); h1 |= (unsigned)hdr[4 + k] << (8 * k); }
  reinterpret_cast<unsigned *>(dg)[0] = h0;
  reinterpret_cast<unsigned *>(dg)[1] = h1;
  valid[i] = a != ACT_NONE;
}

// after pass 3: the cache entries that changed in this batch become the state carried to the next one
__global__ void k_trx_commit(int narfcn, const int *__restrict__ commit, const DfeRec *__restrict__ dfe, TrxState *__restrict__ st) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= narfcn * 8) return;
  const int src = commit[k];
  if (src < 0) return;
  const int a = k >> 3, tn = k & 7;
  const DfeRec r = dfe[src];
  for (int j = 0; j < 7; j++) st[a].w[tn][j] = r.w[j];
  for (int j = 0; j < 5; j++) st[a].b[tn][j] = r.b[j];
  st[a].chan_off[tn] = r.off;
}

struct TrxScratch {                     // device scratch of one pull (caller-owned, sized by trx_scratch_bytes)
  DetRec *det; int *act; double *thr_at; DfeRec *dfe; EqParams *eqp; int *commit;
  int *rach_flag; cf *rach_amp; float *rach_toa; float *rach_soft; EqParams *eqp_r; cf *rach_cs;
};
constexpr int kTrxRachSoftPitch = 160;
// slice_all: the no-equaliser mode demodulates every accepted burst into a float row of its own (n rows instead of nr)
size_t trx_scratch_bytes(long long n, long long nr, int narfcn, bool slice_all) {
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  if (slice_all) nr = n;
  return up(n * sizeof(DetRec)) + up(n * 4) + up(n * 8) + up(n * sizeof(DfeRec)) + up(n * sizeof(EqParams)) + up((size_t)narfcn * 8 * 4) +
         up(nr * 4 + 4) + up(nr * 8 + 8) + up(nr * 4 + 4) + up(nr * kTrxRachSoftPitch * 4 + 4) + up(nr * sizeof(EqParams) + 16) + up(nr * 160 * sizeof(cf) + 16);
}
static TrxScratch trx_carve(void *base, long long n, long long nr, int narfcn, bool slice_all) {
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  if (slice_all) nr = n;
  char *p = reinterpret_cast<char *>(base);
  TrxScratch s;
  s.det = (DetRec *)p; p += up(n * sizeof(DetRec));
  s.act = (int *)p; p += up(n * 4);
  s.thr_at = (double *)p; p += up(n * 8);
  s.dfe = (DfeRec *)p; p += up(n * sizeof(DfeRec));
  s.eqp = (EqParams *)p; p += up(n * sizeof(EqParams));
  s.commit = (int *)p; p += up((size_t)narfcn * 8 * 4);
  s.rach_flag = (int *)p; p += up(nr * 4 + 4);
  s.rach_amp = (cf *)p; p += up(nr * 8 + 8);
  s.rach_toa = (float *)p; p += up(nr * 4 + 4);
  s.rach_soft = (float *)p; p += up(nr * kTrxRachSoftPitch * 4 + 4);
  s.eqp_r = (EqParams *)p; p += up(nr * sizeof(EqParams) + 16);
  s.rach_cs = (cf *)p;
  return s;
}

// The whole pull for one batch.  kind / tsc: one byte per burst (CorrType, midamble code); rach_idx: the nr bursts in
// RACH slots; rach_slot: per burst, its index in rach_idx or -1.  All device pointers.  Returns the launch count.
// pitch > 0: bursts laid out [frame][arfcn][tn] at that pitch; pitch == 0: `bursts` holds narfcn continuous slot
// streams, stream_pitch samples apart (what the RX resampler writes).
int launch_trx_pull(const DevTables *T, TrxState *st, int narfcn, int nframes, int fn0, const cf *bursts, long long pitch,
                    long long stream_pitch, const uint8_t *kind, const uint8_t *tsc, const int *rach_idx, const int *rach_slot,
                    long long nr, void *scratch, int *valid, unsigned char *dgram, int dgram_pitch, cudaStream_t stream,
                    cudaStream_t side, cudaEvent_t *ev, TrxVariant var) {
  const long long n = (long long)nframes * narfcn * 8;
  if (n <= 0) return 0;
  const bool slice_all = var.v52m && !var.need_dfe;
  // The access-burst kernels work on a few percent of the slots and fill a fraction of the GPU: they run on a side
  // stream (when the caller supplies one and four events) next to the normal-burst kernels they do not depend on.
  const bool fork = side != nullptr && ev != nullptr && nr > 0;
  cudaStream_t rs = fork ? side : stream;
  const TrxScratch s = trx_carve(scratch, n, nr, narfcn, slice_all);
  BurstSrc src{bursts, pitch, nullptr, 0, 1};
  if (pitch == 0) { src.narfcn = narfcn; src.arfcn_pitch = stream_pitch; }
  int launches = 0;
  const long long nwarps = (n + 31) / 32;
  NormalOut none{};
  // pass 1
  if (fork) { cudaEventRecord(ev[0], stream); cudaStreamWaitEvent(rs, ev[0], 0); }        // inputs (maps, bursts) are ready
  if (var.v52m)      // the second variant's pass 1: stride-4 energy, windowed midamble search (channel only when it can be used)
    k_detect_52m<true><<<(unsigned)nwarps, 32, detect_52m_smem(var.max_toa), stream>>>(T, src, tsc, n, 3.0F, var.max_toa, var.need_dfe ? 1 : 0, none, kind, s.det);
  else if (nwarps >= kDetWideMin)
    k_detect_design<kDetWarps, true><<<(unsigned)((nwarps + kDetWarps - 1) / kDetWarps), 32 * kDetWarps, detect_smem<kDetWarps>(), stream>>>(T, src, tsc, n, 3.0F, 0.0F, 0.0F, none, nullptr, kind, s.det);
  else
    k_detect_design<1, true><<<(unsigned)nwarps, 32, detect_smem<1>(), stream>>>(T, src, tsc, n, 3.0F, 0.0F, 0.0F, none, nullptr, kind, s.det);
  launches++;
  BurstSrc rsrc = src;
  rsrc.gather = rach_idx;
  if (nr > 0) {
    NormalOut ro{};
    ro.flag = s.rach_flag; ro.amp = s.rach_amp; ro.toa = s.rach_toa;
    k_rach_detect<<<(unsigned)((nr + 31) / 32), 32, kRachRollBytes, rs>>>(T, rsrc, nr, 5.0F, ro, s.eqp_r, s.rach_cs);
    if (fork) { cudaEventRecord(ev[1], rs); cudaStreamWaitEvent(stream, ev[1], 0); }      // the policy needs the RACH flags
    launches++;
  }
  // pass 2
  k_trx_policy<<<(narfcn + 31) / 32, 32, 0, stream>>>(T, st, narfcn, nframes, fn0, s.det, rach_slot, s.rach_flag, s.act, s.thr_at, s.commit,
                                                      slice_all ? 0 : 1);
  if (slice_all) {
    // pass 3 without the equaliser: one demodulateBurst pass over the whole batch (accepted normal AND access bursts),
    // then the datagrams from each burst's own soft row
    k_trx_slice_params<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(n, s.det, s.act, rach_slot, s.rach_amp, s.rach_toa, s.eqp);
    launch_slicer(T, src, n, s.eqp, s.rach_soft, kTrxRachSoftPitch, stream);
    k_trx_datagram<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(T, n, narfcn, fn0, s.det, s.act, rach_slot, s.rach_amp, s.rach_toa,
                                                                        s.rach_soft, kTrxRachSoftPitch, valid, dgram, dgram_pitch, s.rach_soft);
    return launches + 4;
  }
  // pass 3
  k_trx_design<<<(unsigned)((n + 63) / 64), 64, 0, stream>>>(n, s.det, s.act, s.thr_at, s.dfe);
  if (fork) { cudaEventRecord(ev[2], stream); cudaStreamWaitEvent(rs, ev[2], 0); }        // act[] is final
  k_trx_eqparams<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(n, narfcn, s.det, s.act, s.dfe, st, s.eqp);
  if (g_eq_ring) k_equalize_ring<true><<<(unsigned)nwarps, 32, kEqRingBytes, stream>>>(T, src, n, s.eqp, dgram + 8, dgram_pitch, 152);
  else k_equalize_fast<1, true><<<(unsigned)nwarps, 32, equalize_smem<1>(), stream>>>(T, src, n, s.eqp, dgram + 8, dgram_pitch, 152);
  launches += 4;
  if (nr > 0) {
    k_trx_rach_veto<<<(unsigned)((nr + 127) / 128), 128, 0, rs>>>(nr, rach_idx, s.act, s.eqp_r);
    launch_slicer(T, rsrc, nr, s.eqp_r, s.rach_soft, kTrxRachSoftPitch, rs);
    if (fork) { cudaEventRecord(ev[3], rs); cudaStreamWaitEvent(stream, ev[3], 0); }      // the datagrams need the RACH soft bits
    launches += 2;
  }
  k_trx_datagram<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(T, n, narfcn, fn0, s.det, s.act, rach_slot, s.rach_amp, s.rach_toa,
                                                                      s.rach_soft, kTrxRachSoftPitch, valid, dgram, dgram_pitch);
  k_trx_commit<<<(narfcn * 8 + 127) / 128, 128, 0, stream>>>(narfcn, s.commit, s.dfe, st);
  return launches + 2;
}
