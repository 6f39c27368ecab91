/*
 * oracle/sigproc_port.h -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 * Plain-C restatement of the reference burst-DSP path; see sigproc_port.c for the per-function
 * reference citations.  Same entry points as oracle/ref_shim.cpp with prefix port_.
 * Complex vectors are interleaved float (re, im); bits one char each; soft bits one float each.
 */
#ifndef SIGPROC_PORT_H
#define SIGPROC_PORT_H
#ifdef __cplusplus
extern "C" {
#endif

/* ConvType, reference sigProcLib.h:40-47 */
enum { PORT_FULL_SPAN = 0, PORT_OVERLAP_ONLY = 1, PORT_START_ONLY = 2, PORT_WITH_TAIL = 3, PORT_NO_DELAY = 4 };

int   port_setup(int sps);
int   port_get_table(int id, int idx, float *dst, int cap);
float port_sinc(float x);
float port_sin_lookup(float x);
float port_cos_lookup(float x);
int   port_modulate(const char *bits, int nbits, int guard, int sps, float *out, int cap);
void  port_delay_vector(float *v, int n, float delay);
void  port_scale_vector(float *v, int n, int real_only, const float *scale);
int   port_convolve(const float *a, int la, int a_real, const float *b, int lb, int b_real, float *c, int cap, int span);
int   port_correlate(const float *a, int la, int a_real, const float *b, int lb, int b_real, float *c, int cap, int span);
void  port_peak_detect(const float *v, int n, float *peak, float *idx, float *avg);
void  port_interpolate_point(const float *v, int n, float ix, float *out);
int   port_energy_detect(const float *v, int n, unsigned win, float thr, float *avg);
int   port_analyze(const float *burst, int n, int tsc, float thr, int sps, float *amp, float *toa,
                   int request, float *chan, float *off);
int   port_detect_rach(const float *burst, int n, float thr, int sps, float *amp, float *toa);
int   port_design_dfe(const float *chan, int nchan, float snr, int Nf, float *w, float *b);
int   port_equalize(float *burst, int n, float toa, int sps, const float *w, int nw, const float *b, int nb, float *soft);
int   port_demodulate(const float *burst, int n, int sps, const float *amp, float toa, float *soft);
int   port_resample(const float *x, int n, int P, int Q, int lpf, float *out, int cap);

void  port_rx_normal_batch(const float *bursts, int pitch, const int *lens, const unsigned char *tsc, long n,
                           float detect_thr, float energy_thr,
                           int *flags, float *amp, float *toa, float *chan, float *off, float *w, float *b,
                           float *soft, int soft_pitch);
void  port_rx_rach_batch(const float *bursts, int pitch, const int *lens, long n, float detect_thr, int sps,
                         int *flags, float *amp, float *toa, float *soft, int soft_pitch);
void  port_rx_resample_stream(const float *raw, long first_chunk, long nchunks, float *out);
void  port_rx_resample_stream_i16(const short *iq, int flip_iq, long first_chunk, long nchunks, float *out);
void  port_soft_to_wire(const float *soft, int soft_pitch, long n, unsigned char *out);
void  port_tx_resample_stream(const float *in, long first_chunk, long nchunks, short *out);
long  port_modulate_stream(const char *bits148, long nbursts, int tn0, float *out);
void  port_rx_stream_demod(const float *resampled, long first_burst, long nbursts, const unsigned char *tsc,
                           float detect_thr, float energy_thr, int *flags, float *amp, float *toa,
                           float *soft, int soft_pitch);
/* Transceiver::pullRadioVector policy + RX datagram (one state = one ARFCN); see sigproc_port.c */
int   port_expected_corr_type(int chan_type, int fn);
int   port_trx_state_bytes(void);
void  port_trx_init(void *state, int tsc, const int *chan_type, int start_fn);
void  port_trx_pull(void *state, const float *bursts, int pitch, int nframes, int fn0, int *valid,
                    unsigned char *dgram, int dgram_pitch);
/* XCCH block decoder (deinterleave + soft Viterbi + Fire-code syndrome); see sigproc_port.c */
void  port_xcch_decode(const unsigned char *soft, int burst_pitch, long nframes, unsigned char *u, int *ok);
void  port_tch_decode(const unsigned char *soft, int burst_pitch, long nblocks, unsigned char *d, int *good, int *stolen,
                      unsigned char *fu, int *fok);
#ifdef __cplusplus
}
#endif
#endif
