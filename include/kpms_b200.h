/* C-ABI of libkpms_b200.so: hand-written sm_100a kernels for the keypoint-SLDS Gibbs sweep.
 *
 * This is the drop-in boundary below the reference's single call per sweep,
 *   model = resample_func(data, **model, **resample_options)
 *   (/root/reference/keypoint_moseq/fitting.py:25, resample_func bound at :245-248, :396-399, :510-513),
 * i.e. what a ctypes / cffi binding of jax_moseq.models.keypoint_slds.resample_model's
 * sub-samplers would call.  Each entry point names the upstream function it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated; arrays are dense row-major;
 *   - dtype: 0 = float32, 1 = float64 for all `void*` real arrays of the call;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it;
 *   - return 0 on success, a negative argument error or a positive cudaError_t otherwise;
 *     kpms_last_error() returns the message of the calling thread's last failure;
 *   - no allocation happens inside: scratch comes from the caller (`*_workspace_bytes`);
 *   - `*_tape` arguments are nullable: non-null = verification mode (injected standard
 *     normals / uniforms, layouts in DESIGN.md), null = Philox4x32-10 keyed by `seed`;
 *   - every sampler takes its key as `uint64_t seed, const uint64_t* seed_dev`: with seed_dev == NULL the key is
 *     `seed`; otherwise it is `*seed_dev XOR seed`, read on the device at run time, so that a sweep captured in a
 *     CUDA graph draws fresh numbers at every replay (kpms_advance_seed steps the device-resident key);
 *   - N chains (segments), T frames per chain, k keypoints, Dk in {2,3}, d = latent_dim,
 *     L = nlags, n = d*L, K = num_states, Tp = T - L, Tx = T - L + 1.
 */
#ifndef KPMS_B200_H
#define KPMS_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KPMS_F32 0
#define KPMS_F64 1
#define KPMS_GAMMA_ATTEMPTS 6   /* taped Marsaglia-Tsang attempts; tape row = 2*6+1 values */
#define KPMS_VONMISES_ATTEMPTS 8 /* taped Best-Fisher attempts; tape row = 8*3 uniforms */

int kpms_version(void);
/* *seed_dev <- splitmix64 step of *seed_dev (one thread; the per-sweep key schedule of resample_model, which the
 * reference drives with jax.random.split on model["seed"], fitting.py:25). */
int kpms_advance_seed(uint64_t* seed_dev, void* stream);
const char* kpms_last_error(void);
/* launch accounting: number of kernels launched by this library so far; optional per-kernel
 * CUDA-event timing (enable, run, then report "name total_ms count" lines; report synchronises). */
long long kpms_launch_count(void);
void kpms_profile_enable(int on);
int kpms_profile_report(char* buf, size_t cap);

/* Time-parallel execution of the serial recursions (Kalman filter, backward sampler, HMM filter).
 * Each chain is cut into `chunks` pieces that run concurrently from `warmup` steps early; every
 * boundary is checked against its neighbour and chains whose discrepancy exceeds the tolerance
 * (tol32 for float32 calls, tol64 for float64) are re-run sequentially inside the same call.
 * chunks = 0: automatic (fill the device), 1: sequential; a negative / non-positive argument
 * leaves that setting unchanged.  Defaults: 0, 64, 2e-5, 1e-10.
 * After a call the first 16 bytes of that call's workspace hold, as uint32: the largest forward
 * boundary discrepancy (float bits), the number of chains re-run forward, and the same two for
 * the backward recursion. */
void kpms_set_time_chunking(int chunks, int warmup, double tol32, double tol64);
/* The chunk count the library picks for N chains of `len` steps when `slots` chunk tasks are resident at once: the C
 * that minimises ceil(N C / slots) * (len / C + warmup), smallest C within 3 % of the optimum, at most len / (4 warmup);
 * the value forced by kpms_set_time_chunking when there is one.  Host-only (planning / tests). */
int kpms_plan_chunks(int N, int slots, int len, int warmup);

/* ---- discrete states: jax_moseq.models.arhmm.resample_discrete_stateseqs
 *      (utils.autoregression.ar_log_likelihood + utils.distributions.sample_hmm_stateseq);
 *      the forward pass alone is arhmm.marginal_log_likelihood (fitting.py:667-673) and
 *      forward + smooth is arhmm.stateseq_marginals (fitting.py:536-538). */
/* one workspace serves kpms_ar_loglik, kpms_hmm_forward and kpms_hmm_backward_sample of the same (N, T, K, d, L) */
size_t kpms_hmm_workspace_bytes(int dtype, int N, int T, int K, int d, int L);
/* W <- exp(ll - max_k ll), mx (N,ldT) <- max_k ll; masked frames have ll = 0. ldT % 8 == 0.
 * W is an opaque buffer of kpms_hmm_weights_bytes() handed from kpms_ar_loglik to kpms_hmm_forward
 * of the same dtype: (N,K,ldT) time-contiguous in float32, (N,Tp,K rounded up to whole 8-state tiles)
 * state-contiguous in float64 (operand layout of the FP64 tensor-pipe kernels). */
size_t kpms_hmm_weights_bytes(int dtype, int N, int T, int K, int L);
int kpms_ar_loglik(int dtype, const void* x, const int32_t* mask, const void* Ab, const void* Q, int N,
                   int T, int d, int L, int K, int ldT, void* W, void* mx, void* ws, void* stream);
/* filt (N,Tp,ldK), ldK = K rounded up to 4; logZ (N) double = per-chain log normaliser. */
int kpms_hmm_forward(int dtype, const void* W, const void* mx, const void* pi, int N, int K, int Tp,
                     int ldT, void* filt, double* logZ, void* ws, int d, int L, void* stream);
/* z (N,Tp) int32; u_tape (N,Tp) uniforms or NULL, in which case u_scratch (N,Tp) receives Philox uniforms.
 * Exact and parallel in time: with the uniforms fixed the sampler is a deterministic map z_{t+1} -> z_t,
 * so time chunks started from a guess merge with the true path; every chunk boundary is compared with
 * the label actually produced above it and un-merged stretches are re-walked (workspace words 2, 3 =
 * mismatched boundaries, re-walked steps).  `ws` must be the workspace kpms_ar_loglik ran on. */
int kpms_hmm_backward_sample(int dtype, const void* filt, const void* pi, const void* u_tape, void* u_scratch,
                             uint64_t seed, const uint64_t* seed_dev, int N, int K, int Tp, int32_t* z, void* ws,
                             int d, int L, void* stream);
/* marg (N,Tp,K) smoothed marginals from the stored filter. */
int kpms_hmm_smooth(int dtype, const void* filt, const void* pi, int N, int K, int Tp, void* marg,
                    void* stream);

/* ---- continuous states: jax_moseq.models.keypoint_slds.resample_continuous_stateseqs
 *      (-> slds.resample_continuous_stateseqs -> utils.kalman.kalman_sample).
 *      Ct (k*Dk, d+1) = (Gamma kron I) Cd; w_tape (N,Tx,n) normals or NULL; x (N,T,d) out. */
size_t kpms_kalman_workspace_bytes(int dtype, int N, int T, int d, int L, int K);
/* info_ready != 0: the per-frame observation records of this workspace were already computed by kpms_kalman_obs_info
 * (same Y, v, h, s, Ct, sigmasq), e.g. on another stream beside the discrete-state kernels. */
int kpms_kalman_sample(int dtype, const void* Y, const int32_t* mask, const void* v, const void* h,
                       const void* s, const int32_t* z, const void* Ct, const void* sigmasq,
                       const void* Ab, const void* Q, double jitter, const void* w_tape, uint64_t seed,
                       const uint64_t* seed_dev, int N, int T, int k, int Dk, int d, int L, int K, int info_ready,
                       void* x, void* ws, void* stream);
/* first stage of the sampler alone: un-rotated, centred keypoints -> chol(C' R_t^-1 C) and C' R_t^-1 (y_t - d) per
 * frame, written into the workspace (it depends on s, v, h but not on z or the AR parameters). */
int kpms_kalman_obs_info(int dtype, const void* Y, const int32_t* mask, const void* v, const void* h, const void* s,
                         const void* Ct, const void* sigmasq, int N, int T, int k, int Dk, int d, int L, int K,
                         void* ws, void* stream);
/* (latent_dim, nlags) pairs the kernels were compiled for (io.py:72-83 leaves both to the user's config): writes up
 * to `cap` pairs as d0, L0, d1, L1, ... and returns how many exist.  Any other pair fails with status -3 in
 * kpms_ar_loglik / kpms_ar_suffstats / kpms_kalman_sample. */
int kpms_supported_dims(int* pairs, int cap);

/* ---- per-keypoint noise scales: jax_moseq.models.keypoint_slds.resample_scales.
 *      g_tape (N,T,k,13) gamma tape or NULL; noise_prior, s_out (N,T,k). */
int kpms_resample_scales(int dtype, const void* Y, const void* x, const void* v, const void* h,
                         const void* Ct, const void* sigmasq, const void* noise_prior, double nu_s,
                         const void* g_tape, uint64_t seed, const uint64_t* seed_dev, int N, int T, int k, int Dk,
                         int d, void* s_out, void* stream);

/* ---- heading + centroid: keypoint_slds.resample_heading followed by keypoint_slds.resample_location
 *      (which reuses utils.kalman.kalman_sample with identity dynamics).  h_out (N,T), v_out (N,T,Dk);
 *      u_tape (N,T,8,3) uniforms or NULL; w_tape (N,T,Dk) normals or NULL; fix_heading copies h_in. */
size_t kpms_heading_location_workspace_bytes(int dtype, int N, int T, int Dk);
int kpms_resample_heading_location(int dtype, const void* Y, const int32_t* mask, const void* x,
                                   const void* v_in, const void* h_in, const void* s, const void* Ct,
                                   const void* sigmasq, double sigmasq_loc, int fix_heading,
                                   const void* u_tape, const void* w_tape, uint64_t seed,
                                   const uint64_t* seed_dev, int N, int T, int k, int Dk, int d, void* h_out,
                                   void* v_out, void* ws, void* stream);

/* ---- sufficient statistics (the only data that crosses GPUs; all-reduce these)
 *      counts (K,K) int32: utils.transitions.count_transitions;
 *      gram (K,F,F) double, F = n+d+1, features [x_{t-L..t-1} | x_t | 1]: the masked einsums of
 *      arhmm.gibbs._resample_regression_params;
 *      obsvar out (k+1) double: sum_t mask*sqerr/s per keypoint, then the valid-frame count. */
int kpms_transition_counts(const int32_t* z, const int32_t* mask, int N, int T, int L, int K,
                           int32_t* counts, void* stream);
size_t kpms_ar_suffstats_workspace_bytes(int N, int T, int d, int L, int K);
int kpms_ar_suffstats(int dtype, const void* x, const int32_t* z, const int32_t* mask, int N, int T, int d,
                      int L, int K, double* gram, void* ws, void* stream);
size_t kpms_obsvar_workspace_bytes(int N, int T, int k);
int kpms_obsvar_suffstats(int dtype, const void* Y, const int32_t* mask, const void* x, const void* v,
                          const void* h, const void* s, const void* Ct, int N, int T, int k, int Dk, int d,
                          double* out, void* ws, void* stream);

/* ---- parameter draws from the (all-reduced) statistics, always float64
 *      arhmm.resample_ar_params / utils.distributions.sample_mniw:
 *        tapes w_G (K,d,n+1) normals, w_B (K,d,d) normals (strict lower used), g_chi (K,d,13) gamma tape;
 *      utils.transitions.resample_hdp_transitions:
 *        u_crp, u_bin flat uniforms (length >= number of counted transitions), g_beta (K,13), g_pi (K,K,13);
 *      keypoint_slds.resample_obs_variance: g_sig (k,13). */
int kpms_resample_ar_params(const double* gram, const double* K_0, const double* M_0, const double* S_0,
                            double nu_0, const double* w_G, const double* w_B, const double* g_chi,
                            uint64_t seed, const uint64_t* seed_dev, int K, int d, int L, double* Ab, double* Q,
                            void* stream);
size_t kpms_transitions_workspace_bytes(int K);
int kpms_resample_hdp_transitions(const int32_t* counts, const double* betas_in, double alpha, double kappa,
                                  double gamma, const double* u_crp, const double* u_bin, const double* g_beta,
                                  const double* g_pi, uint64_t seed, const uint64_t* seed_dev, int K, double* betas_out,
                                  double* pi, void* ws, void* stream);
int kpms_resample_obs_variance(const double* stats, double nu_sigma, double sigmasq_0, int Dk,
                               const double* g_sig, uint64_t seed, const uint64_t* seed_dev, int k, double* sigmasq,
                               void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KPMS_B200_H */
