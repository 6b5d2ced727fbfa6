/*
 * opticomm_b200.h — C-ABI of the B200-native (sm_100a) hot path for OptiCommPy.
 *
 * Drop-in boundary.  The reference has no FFI: its boundary is a set of
 * Python call signatures taking numpy arrays + a `parameters` attribute bag.  Each entry
 * point below is what a ctypes binding for one of those calls binds to; the citation is
 * the reference function it replaces (paths relative to the OptiCommPy v0.11.0 tree).
 *
 * Conventions
 *   - plain C types only; `void*` device pointers are owned by the caller (torch tensors
 *     used as device-memory containers); `stream` is a cudaStream_t passed as void*.
 *   - every call returns 0 on success, non-zero otherwise; ocb_last_error() returns a
 *     thread-local message for the last failure.
 *   - complex64 samples are (re, im) float pairs; complex128 are (re, im) double pairs.
 *   - field layout inside the library is PLANAR: rows[2K][N]; row p (p<K) is the x-pol of
 *     pol-pair p, row K+p its y-pol (reference: Ei[:, 0::2].T / Ei[:, 1::2].T,
 *     optic/models/channels.py:364-365).
 *   - "_host" variants take HOST pointers and perform H2D / D2H themselves (this is the
 *     call the end-to-end number in bench.py goes through).
 *   - threads: the library keeps no global mutable state apart from a mutex-protected cache of
 *     cuFFT handles keyed by (device, stream, geometry).  Calls from different host threads are
 *     independent as long as every thread uses its own stream and its own ocb_ssfm_plan (a plan
 *     owns one workspace and one convergence mailbox): this is how several independent waveforms
 *     are kept in flight on one GPU (opticommpy_b200.sharding.run_concurrent).
 */
#ifndef OPTICOMM_B200_H
#define OPTICOMM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCB_ABI_VERSION 1

/* dtype tags for host/device sample buffers */
#define OCB_C64 0  /* complex64  */
#define OCB_C128 1 /* complex128 */

/* amplifier modes (optic/models/channels.py:443-451, optic/dsp/equalization.py:1090-1095) */
#define OCB_AMP_NONE 0
#define OCB_AMP_IDEAL 1
#define OCB_AMP_EDFA 2

/* noise source for OCB_AMP_EDFA */
#define OCB_NOISE_INJECTED 0 /* caller supplies one (rows_per_pol, N) complex64 buffer that is
                                added to x and y of every span — exactly what the reference CPU
                                path does when param.seed is set (channels.py:356, 443-445) */
#define OCB_NOISE_PHILOX 1   /* on-device Philox4x32-10 + Box-Muller, independent per pol/span */

/* equalizer algorithms (optic/dsp/equalization.py:477-510) */
#define OCB_ALG_CMA 0
#define OCB_ALG_RDE 1
#define OCB_ALG_NLMS 2
#define OCB_ALG_DDLMS 3
#define OCB_ALG_DARDE 4
#define OCB_ALG_STATIC 5

typedef struct ocb_ssfm_plan ocb_ssfm_plan; /* opaque: cuFFT plans + workspace partition */

/* ---- library ------------------------------------------------------------------------ */
int ocb_abi_version(void);
const char* ocb_last_error(void);
/* Number of kernels launched by this library in the calling thread since the last reset
 * (bench.py's "gpu_launches" claim is read from here). */
int64_t ocb_launch_count(void);
void ocb_launch_count_reset(void);

/* ---- SSFM plan lifecycle -------------------------------------------------------------
 * One plan per (device, N, rows).  rows = 2K for manakovSSF/manakovDBP, 1 for ssfm.
 * The plan owns cuFFT handles; all device memory is a caller-provided workspace.       */
int ocb_ssfm_plan_create(int64_t N, int rows, ocb_ssfm_plan** out);
int64_t ocb_ssfm_plan_workspace_bytes(const ocb_ssfm_plan* plan);
int ocb_ssfm_plan_bind_workspace(ocb_ssfm_plan* plan, void* dev_ptr, int64_t bytes);
int ocb_ssfm_plan_destroy(ocb_ssfm_plan* plan);
/* Transform engine of a plan.  AUTO picks the fused four-step kernels (own FFT passes with the
 * linear operator / Kerr rotation / convergence sums fused in) when N is a power of two in
 * 2^16..2^20 and the plan holds one pol-pair (rows = 2) or one row; otherwise the cuFFT-driven
 * engine (any N, any K).  Both engines implement the same reference loop. */
#define OCB_ENGINE_AUTO 0
#define OCB_ENGINE_CUFFT 1
#define OCB_ENGINE_FUSED 2
int ocb_ssfm_plan_set_engine(ocb_ssfm_plan* plan, int engine);
int ocb_ssfm_plan_engine(const ocb_ssfm_plan* plan); /* engine a run would use now */
/* In-situ kernel timing with CUDA events on the launching stream (bench.py's roofline leg).
 * kinds: 0 = fused nonlinear iteration pass, 1 = nonlinear first pass, 2 = linear half step
 * (fft + multiply + ifft).  profile_read: out6[2k] = summed ms, out6[2k+1] = launches, then reset. */
int ocb_ssfm_plan_profile(ocb_ssfm_plan* plan, int enable);
int ocb_ssfm_plan_profile_read(ocb_ssfm_plan* plan, double* out6);
/* Tuning/measurement aid (no reference counterpart): launch one pass kernel of the fused engine `reps` times
 * back to back on the plan's own (L2-resident) buffers, exactly as the step loop launches it, and return the
 * average duration in microseconds (CUDA events on `stream`).  which: 0 = frequency pass (k_freq),
 * 1 = first time pass (TM_FIRST), 2 = iteration time pass (TM_ITER), 3 = predicted-last iteration (TM_ITERF),
 * 4 = forward time pass (TM_FWD).  The field buffers hold garbage afterwards.                          */
int ocb_ssfm_plan_pass_time(ocb_ssfm_plan* plan, int which, int reps, double* avg_us, void* stream);

/* ---- layout conversion ---------------------------------------------------------------
 * (N, C) interleaved-column host/device array <-> planar rows[C][N] complex64.
 * column 2p -> row p, column 2p+1 -> row C/2+p when pairs != 0; identity order otherwise.
 * Replaces: Ei[:, 0::2].T / Ech[:, 0::2] = Ech_x.T (channels.py:364-365, 461-463).      */
int ocb_pack_fields(const void* src_dev, int src_dtype, int64_t N, int C, int pairs,
                    void* rows_dev, void* stream);
int ocb_unpack_fields(const void* rows_dev, int64_t N, int C, int pairs, void* dst_dev,
                      int dst_dtype, void* stream);

/* complex64 <-> complex128 conversion of n samples on the device (host arrays are uploaded raw and
 * converted here: numpy's complex astype is the slowest part of the host shims otherwise). */
int ocb_cast_complex(const void* src_dev, int src_dtype, void* dst_dev, int dst_dtype, int64_t n,
                     void* stream);

/* ---- Manakov SSFM / DBP ---------------------------------------------------------------
 * Replaces optic.models.channels.manakovSSF (channels.py:252-468; direction=+1) and
 * optic.dsp.equalization.manakovDBP (equalization.py:976-1173; direction=-1).          */
typedef struct ocb_manakov_params {
    double alpha_lin;        /* α [1/km]  = alpha_dB/(10 log10 e)      channels.py:346 */
    double beta2;            /* β2 [s^2/km]                            channels.py:347 */
    double gamma;            /* γ [1/W/km]                             channels.py:348 */
    double Fs;               /* sampling rate [Hz]                                     */
    double Lspan;            /* span length [km]                                       */
    double hz;               /* fixed step [km]                        channels.py:398-403 */
    double maxNlinPhaseRot;  /* adaptive-step phase budget [rad]       channels.py:392-397 */
    double tol;              /* fixed-point tolerance                  channels.py:429 */
    int32_t n_spans;
    int32_t maxIter;
    int32_t nlprMethod;      /* 1 = adaptive step */
    int32_t direction;       /* +1 forward (SSF), -1 backward (DBP) */
    int32_t amp_mode;        /* OCB_AMP_* */
    int32_t noise_mode;      /* OCB_NOISE_* (only for OCB_AMP_EDFA, forward) */
    double edfa_gain_lin;    /* G_lin = 10^(alpha*Lspan/10)            devices.py:713 */
    double edfa_noise_var;   /* p_noise = N_ase*Fs                     devices.py:721-722 */
    uint64_t seed;           /* Philox key */
    int32_t n_save;          /* number of span snapshots requested (0 = final field only) */
    int32_t reserved;
} ocb_manakov_params;

typedef struct ocb_manakov_stats {
    int64_t steps;         /* executed SSFM loop steps over all spans (channels.py:387) */
    int64_t iterations;    /* executed fixed-point iterations over all steps (channels.py:413) */
    int64_t nonconverged;  /* steps that hit maxIter without lim < tol (channels.py:431-434) */
    double last_lim;
    double z_last_step;    /* size of the last executed step [km] */
} ocb_manakov_stats;

/* rows_inout: planar (2K, N) complex64, propagated in place.
 * noise_dev : (K, N) complex64 or NULL (OCB_NOISE_INJECTED only).
 * save_spans: n_save ascending 1-based span indices or NULL; snapshots are written to
 *             save_dev as n_save consecutive planar (2K, N) blocks (channels.py:453-456). */
int ocb_manakov_run(ocb_ssfm_plan* plan, void* rows_inout, const ocb_manakov_params* prm,
                    const void* noise_dev, const int32_t* save_spans, void* save_dev,
                    ocb_manakov_stats* stats, void* stream);

/* Host-buffer variant: Ei_host is the reference-layout (N, 2K) array (complex64/128), the
 * result is written to Eo_host with the same layout and out_dtype (H2D, pack, run, unpack,
 * D2H inside).  noise_host: (K, N) complex64 or NULL.  With n_save > 0, Eo_host receives n_save
 * consecutive (N, 2K) blocks (one per snapshot); the Python shim interleaves them column-wise
 * into the reference's (N, 2K*n_save) layout. */
int ocb_manakov_run_host(ocb_ssfm_plan* plan, const void* Ei_host, int in_dtype, void* Eo_host,
                         int out_dtype, const ocb_manakov_params* prm, const void* noise_host,
                         const int32_t* save_spans, ocb_manakov_stats* stats, void* stream);

/* One fixed-point pass of the nonlinear step on caller buffers (unit parity + ncu target).
 * Replaces channels.py:414-417 (rotation) + :424 (convergence sums) + :436 (phase update):
 *   out   = Ehd * exp(j*dir*hz*(8/9)γ(Pch + |Ec_x|^2 + |Ec_y|^2)/2)
 *   sums  = { Σ|Efd-Ec|^2, Σ|Ec|^2, max(|Efd_x|^2+|Efd_y|^2) }   (3 doubles, device)
 * All fields planar (2K, N) complex64; Pch (K, N) float32.  If Efd == NULL the pass is the
 * first one of a step: Ec is the step-start field, Pch is WRITTEN (|Ec_x|^2+|Ec_y|^2).     */
int ocb_manakov_nl_pass(const void* Ehd, const void* Efd, const void* Ec, void* Pch, void* out,
                        void* sums3_dev, int64_t N, int K, double gamma, double hz, int direction,
                        void* stream);

/* ---- scalar NLSE SSFM -----------------------------------------------------------------
 * Replaces optic.models.channels.ssfm (channels.py:112-249).  rows = 1 plan.            */
typedef struct ocb_nlse_params {
    double alpha_lin, beta2, gamma, Fs, hz;
    int32_t n_spans, n_steps; /* Nspans, Nsteps = floor(Lspan/hz)   channels.py:205-206 */
    int32_t amp_mode, noise_mode;
    double edfa_gain_lin, edfa_noise_var;
    uint64_t seed;
} ocb_nlse_params;
int ocb_nlse_run(ocb_ssfm_plan* plan, void* row_inout, const ocb_nlse_params* prm,
                 const void* noise_dev, void* stream);
int ocb_nlse_run_host(ocb_ssfm_plan* plan, const void* Ei_host, int in_dtype, void* Eo_host,
                      int out_dtype, const ocb_nlse_params* prm, const void* noise_host,
                      void* stream);

/* ---- EDFA ------------------------------------------------------------------------------
 * Replaces optic.models.devices.edfa (devices.py:671-726) on planar rows: E = E*sqrt(G)+n. */
int ocb_edfa_apply(void* rows_inout, int rows, int64_t N, double gain_lin, double noise_var,
                   int noise_mode, const void* noise_dev, int noise_rows, uint64_t seed,
                   uint64_t stream_id, void* stream);

/* ---- EDC (overlap-save CD compensation) -------------------------------------------------
 * Replaces optic.dsp.equalization.edc (equalization.py:36-122) ->
 * optic.dsp.core.blockwiseFFTConv(freqDomainFilter=True) (core.py:973-1046).
 * h_taps: the K time-domain taps fftshift(ifft(H)) (host computes them in float64 and passes
 * them as complex64 device array).  x: planar (nModes, L) complex64.  y: same shape.
 * The result equals conv(x, h)[D : D+L], D=(K-1)//2 (core.py:1004, 1044).                */
int64_t ocb_edc_workspace_bytes(int64_t L, int nModes, int K);
int ocb_edc_run(const void* x_rows, void* y_rows, int64_t L, int nModes, const void* h_taps,
                int K, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- N x N adaptive MIMO equalizer --------------------------------------------------------
 * Replaces optic.dsp.equalization.coreAdaptEq (equalization.py:354-516) and the tap-update
 * kernels cmaUp/rdeUp/nlmsUp/ddlmsUp/dardeUp (:520-973) for a batch of independent streams,
 * one persistent warp per stream.  All strides are in ELEMENTS of the respective array, so a
 * training stage is a pointer offset into the full buffers (equalization.py:276-293).
 *   x      : stream s starts at x + s*x_stream_stride; (nSamp, nModes) complex64 samples available,
 *            already zero-padded like equalization.py:227-231
 *   ref    : stream s at ref + s*ref_stream_stride; (L, nModes) complex64 (NLMS / DA-RDE) or NULL
 *   H, Hwl : (nStreams, nModes^2, nTaps) complex64, updated in place (Hwl NULL unless runWL)
 *   y      : stream s at y + s*y_stream_stride; (L, nModes) complex64
 *   errSq  : stream s, mode m at errSq + s*err_stream_stride + m*err_mode_stride; L float32
 *   Hiter  : NULL, or (nStreams, L, nModes^2, nTaps) complex64 tap history (storeCoeff, :511-512)
 *   constSymb (M) complex64 ; radii (nR) float32 ascending (np.unique(|c|), :456) ; Rcma (:453)  */
int ocb_mimo_eq_run(const void* x, const void* ref, void* H, void* Hwl, void* y, void* errSq,
                    void* Hiter, int nStreams, int64_t nSamp, int64_t x_stream_stride,
                    int64_t ref_stream_stride, int64_t y_stream_stride, int64_t err_stream_stride,
                    int64_t err_mode_stride, int64_t L, int nModes, int nTaps, int SpS, int alg,
                    float mu, const void* constSymb, int M, const void* radii, int nR, float Rcma,
                    int runWL, void* stream);

/* ocb_mimo_eq_rls_run: one RLS ('rls', decision_directed = 0, error against ref) or DD-RLS ('dd-rls',
 * decision_directed = 1) training stage.  Replaces coreAdaptEq with rlsUp / ddrlsUp
 * (optic/dsp/equalization.py:354-516, 576-644, 712-785).  Same buffer conventions as ocb_mimo_eq_run; the
 * inverse correlation matrices start from the identity in every call (like the reference's 'rls' branch,
 * :447-451) and live on the device only.  nTaps <= 64 (one matrix row per lane up to 32 taps, two beyond;
 * examples/test_WDM_transmission.ipynb uses 35).  'rls' is pinned to reference golden vectors (nTaps = 11 and 35);
 * 'dd-rls' is pinned to the CPU oracle ONLY: the reference initialises the inverse correlation matrix for
 * alg == 'rls' alone (:447-451), so its own 'dd-rls' stage runs on an uninitialised matrix and cannot produce golden
 * vectors.  workspace: the gain vectors Y_N(s), ocb_mimo_eq_rls_workspace_bytes(nStreams, nModes, L, nTaps) bytes.  */
int64_t ocb_mimo_eq_rls_workspace_bytes(int nStreams, int nModes, int64_t L, int nTaps);
int ocb_mimo_eq_rls_run(const void* x, const void* ref, void* H, void* y, void* errSq, void* Hiter,
                        int nStreams, int64_t nSamp, int64_t x_stream_stride, int64_t ref_stream_stride,
                        int64_t y_stream_stride, int64_t err_stream_stride, int64_t err_mode_stride,
                        int64_t L, int nModes, int nTaps, int SpS, int decision_directed, float lambda,
                        const void* constSymb, int M, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- blind phase search ----------------------------------------------------------------------
 * Replaces optic.dsp.carrierRecovery.bps (carrierRecovery.py:172-223).
 *   x     : (L, nModes) complex128 (double pairs) device array, modes interleaved as in the reference
 *   constSymb : (M) complex128 ; B test phases b*(pi/2)/B ; Nhalf = N of the 2N+1 window
 *   idx_out : (L, nModes) int32 argmin index ; phase_out : (L, nModes) float64 = testPhases[idx] */
int ocb_bps_run(const void* x, int64_t L, int nModes, const void* constSymb, int M, int B, int Nhalf,
                void* idx_out, void* phase_out, void* stream);

/* ---- carrier phase recovery wrapper -----------------------------------------------------------------
 * Replaces optic.dsp.carrierRecovery.cpr with alg='bps' (carrierRecovery.py:110-169) in one call, all on
 * the device and in float64 like the reference: optional fourthPowerFOE (:333-371) + pnorm, bps (:138),
 * unwrap(4*phase)/4 (:154), pnorm(x * exp(j*phase)) (:162).
 *   x_dev : (L, nModes) complex64/128 ; constSymb : (M) complex128 (power-normalised constellation)
 *   y_out : (L, nModes) complex128 ; phase_out : (L, nModes) float64 unwrapped phases
 *   fo_host : nModes estimated frequency offsets [Hz] (host array, may be NULL)                     */
int64_t ocb_cpr_workspace_bytes(int64_t L, int nModes);
int ocb_cpr_bps_run(const void* x_dev, int x_dtype, int64_t L, int nModes, const void* constSymb, int M,
                    int B, int Nhalf, int runFOE, double Fs, int foeM, void* y_out, void* phase_out,
                    double* fo_host, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- hard decisions and Monte-Carlo error counting ------------------------------------------------------
 * ocb_min_euclid replaces optic.comm.modulation.minEuclid (modulation.py:271-299) and, with bits_out, the
 * demodulateGray/demap pair (:369-408, :303-333): idx = argmin_c |x - constSymb[c]| (first index on ties),
 * bits = binary expansion of idx, most significant bit first, log2(M) per symbol.
 *   x_dev : n complex64/128 ; constSymb : (M) complex128 ; idx_out : n int64 or NULL ; bits_out : n*log2(M) int64 or NULL */
int ocb_min_euclid(const void* x_dev, int x_dtype, int64_t n, const void* constSymb, int M, int64_t* idx_out,
                   int64_t* bits_out, void* stream);

/* ocb_ber_count replaces optic.comm.metrics.fastBERcalc (metrics.py:110-195) for (L, nModes) interleaved
 * rx/tx columns on the device: rotate != 0 applies the phase-ambiguity correction mean(tx/rx) (:176-179, 'qam'
 * and 'psk'), both are power-normalised (:181-182), SNR[dB] = 10 log10(P(tx)/P(rx - tx)) (:185), decisions on
 * sqrtEs * x against constSymb (complex128, Gray order), BER/SER from the index XOR (:187-192).
 *   ber/ser/snr_host : nModes doubles each (host, may be NULL) ; counts_host : [2][nModes] bit / symbol errors
 *   (host, may be NULL).  Synchronises the stream before returning.                                      */
int64_t ocb_ber_workspace_bytes(int nModes);
int ocb_ber_count(const void* rx_dev, const void* tx_dev, int dtype, int64_t L, int nModes, const void* constSymb,
                  int M, int rotate, double sqrtEs, double* ber_host, double* ser_host, double* snr_host,
                  int64_t* counts_host, void* workspace, int64_t workspace_bytes, void* stream);

/* ocb_pnorm_run replaces optic.dsp.core.pnorm (optic/dsp/core.py:702-717): x / sqrt(mean |x|^2) over the whole
 * array, in place.  x_dev: n complex128 ; workspace: at least 4096 doubles.                                      */
int ocb_pnorm_run(void* x_dev, int64_t n, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- Rx front-end glue --------------------------------------
 * decimate: maximum-variance sampling instant per mode + downsampling.
 * Replaces: optic.dsp.core.decimate (optic/dsp/core.py:435-491).
 * x_rows: planar [nModes][N] complex64; y_rows: [nModes][ceil(N/decFactor)]; delays_dev: int32[nModes].
 * (firFilter, optic/dsp/core.py:87-125, is fftconvolve(x, h, 'same') per mode = ocb_edc_run with taps h.) */
int64_t ocb_decimate_workspace_bytes(int nModes, int SpSin);
int ocb_decimate_run(const void* x_rows, void* y_rows, int64_t N, int nModes, int SpSin, int decFactor,
                     void* delays_dev, void* workspace, int64_t workspace_bytes, void* stream);

/* Polarisation-multiplexed coherent front end with ideal photodiodes, one pass over the samples.
 * Replaces: optic.models.devices.pdmCoherentReceiver (optic/models/devices.py:574-668) with paramPD.ideal = True and
 * without polarisation delay / IQ skew (the host mirror composes those from the FIR path): pbs rotation of the signal
 * (:223-262), PDL, 45-degree LO split, opticalHybrid2x4 (:447-503), balancedPD (:402-444) with photodiode R |E|^2
 * (:313-318), iqMixing amplitude / phase imbalance (optic/dsp/core.py:951-959; iq_k = k1x, k2x, k1y, k2y as (re, im)).
 *   Es_rows : planar [2][N] complex64 (x, y) ; S_rows : [2][N] complex64 out
 *   Elo     : [N] complex64 LO field, or NULL for a noiseless CW LO sqrt(lo_power_w) exp(j 2 pi lo_freq_shift n / Fs)
 *             generated on the fly — the channel down-shift of a WDM receiver (basicLaserModel with lw = 0, RIN = 0,
 *             devices.py:729-791).                                                                                  */
int ocb_pdm_frontend_run(const void* Es_rows, const void* Elo, void* S_rows, int64_t N, double polRotation,
                         double pdl_dB, double R, double lo_power_w, double lo_freq_shift, double Fs,
                         const double* iq_k, void* stream);
/* rows[r][n] *= exp(-j 2 pi freq n / Fs): stand-alone frequency down-shift of planar complex64 rows, in place. */
int ocb_freq_shift_run(void* rows, int nRows, int64_t N, double freq, double Fs, void* stream);

/* symbolSync building blocks (optic/dsp/core.py:552-675; finddelay :678-698).  The decisions (column swap, pi/2
 * rotation, conjugation, delay) are scalars taken on the host from the correlation peaks; everything of length L runs
 * on the device.
 *   ocb_sync_sequence_run : column r of an (L, nCols) complex128 array -> real row out[r][L] (double):
 *                           kind 0: |z| - mean|z| (:605-608), 1: Re z, 2: Im z (:628-629)
 *   ocb_xcorr_peak_run    : for every pair (a_i, b_j) of real rows, the first index k maximising |c[k]| of
 *                           c = scipy.signal.correlate(a_i, b_j, 'full') and the value c[k]; results in host arrays
 *                           indexed i*nB + j.  Synchronises the stream.
 *   ocb_sync_apply_run    : out[n][k] = conj?(rot[k] * tx[(n + delay[k]) mod L][swap[k]])  (:648-666), complex128   */
int ocb_sync_sequence_run(const void* z_dev, int nCols, int64_t L, int kind, void* out_rows, void* stream);
int64_t ocb_xcorr_workspace_bytes(int nA, int64_t La, int nB, int64_t Lb);
int ocb_xcorr_peak_run(const void* a_rows, int nA, int64_t La, const void* b_rows, int nB, int64_t Lb,
                       int64_t* peak_idx_host, double* peak_val_host, void* workspace, int64_t workspace_bytes,
                       void* stream);
int ocb_sync_apply_run(const void* tx_dev, void* out_dev, int64_t L, int nCols, const int32_t* swap_dev,
                       const void* rot_dev, const int32_t* conj_dev, const int64_t* delay_dev, void* stream);

/* ---- WDM transmitter (input generation at scale) -------------------------------
 * Replaces the per-sample work of optic.models.tx.simpleWDMTx (optic/models/tx.py:42-228); the symbol draw (numpy's
 * legacy generator, optic/comm/sources.py:167-211) and the laser phase-noise walk stay on the host.
 *   ocb_upsample_run       : rows_out[r][SpS*k] = sym_rows[r][k], zeros in between (optic/dsp/core.py:395-432);
 *                            sym_rows planar [nRows][nSym] complex64, rows_out [nRows][nSym*SpS].  The pulse-shaping
 *                            filter is then ocb_edc_run with the pulse taps (firFilter, core.py:87-125).
 *   ocb_wdm_tx_combine_run : shaped_rows planar [nCh*nPol][N] complex64 (row = ch*nPol + mode), per row
 *                            s / max|s| (tx.py:196) -> iqm(LO, mzmScale * s) (optic/models/devices.py:147-216, calcMZM /
 *                            calcPM optic/dsp/core.py:1075-1130) -> sqrt(P_ch / nPol) * pnorm (tx.py:206, core.py:702-717)
 *                            -> freqShift by ch_freq_hz[ch] (core.py:1050-1072) -> summed over the channels in ascending
 *                            order into out_rows planar [nPol][N] complex64 (tx.py:208).
 *                            lo_rows: [nCh][N] complex64 LO fields exp(j phi_pn) or NULL for an ideal laser (linewidth 0);
 *                            ch_power_w / ch_freq_hz: HOST arrays of nCh doubles.                                     */
typedef struct {
    double mzmScale; /* Vrf / Vpi scale of the driving signal (tx.py:64) */
    double Vpi, VbI, VbQ, Vphi, ERI, ERQ; /* iqm parameters, defaults 2, -2, -2, 1, 60, 60 (devices.py:181-186) */
} ocb_wdm_tx_params;
int ocb_upsample_run(const void* sym_rows, int nRows, int64_t nSym, int SpS, void* rows_out, void* stream);
int64_t ocb_wdm_tx_workspace_bytes(int nCh, int nPol);
int ocb_wdm_tx_combine_run(const void* shaped_rows, const void* lo_rows, int nCh, int nPol, int64_t N,
                           const ocb_wdm_tx_params* q, const double* ch_power_w, const double* ch_freq_hz, double Fs,
                           void* out_rows, void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OPTICOMM_B200_H */
