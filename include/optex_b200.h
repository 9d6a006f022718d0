/*
 * optex_b200.h - C-ABI of the B200-native sliced-optimal-transport step.
 *
 * This is the drop-in boundary for the hot path of JCBrouwer/OptimalTextures
 * (reference commit f201bfa).  Every entry point replaces one reference
 * function; the citation after "replaces:" is file:line in the reference tree.
 *
 * Conventions
 *   - plain C, no torch types; device pointers are raw CUDA pointers (fp32),
 *     `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - every function ENQUEUES on `stream` and returns without synchronising,
 *     except the *_host entry points (host buffers in, host buffers out) which
 *     synchronise `stream` before returning.
 *   - return value: OPTEX_OK (0) or an OPTEX_E* code; optex_last_error() gives
 *     a thread-local message.  Nothing aborts the process, nothing falls back
 *     to the CPU: on a non-sm_100 device every compute call returns
 *     OPTEX_EDEVICE.
 *   - feature blocks are "NHWC flattened": row-major [n, c] with c contiguous
 *     (n = b*h*w), exactly the memory the reference hands to `@ rotation`
 *     (optex.py:170-171).  "channel-major" means row-major [c, n].
 *   - inputs are never written; outputs never alias inputs.
 */
#ifndef OPTEX_B200_H
#define OPTEX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OPTEX_ABI_VERSION 1

/* status codes */
#define OPTEX_OK 0
#define OPTEX_EINVAL 1     /* bad argument (shape, NULL pointer, unknown mode) */
#define OPTEX_EDEVICE 2    /* current device is not sm_100 (no fallback exists)  */
#define OPTEX_ECUDA 3      /* a CUDA runtime call or kernel launch failed        */
#define OPTEX_EWORKSPACE 4 /* workspace pointer NULL or too small                */
#define OPTEX_ESIZE 5      /* size outside what the kernels support              */
#define OPTEX_EUNSUPPORTED 6 /* the request needs something this build / process lacks (e.g. NCCL not loadable,
                                a mode without a sharded form)                    */

/* hist modes: histmatch.py:5 `mode`; OPTEX_MODE_SORT is the north-star's exact
 * 1-D OT (not in the reference, defined by oracle/sort_oracle.py).            */
#define OPTEX_MODE_CHOL 0
#define OPTEX_MODE_PCA 1
#define OPTEX_MODE_SYM 2
#define OPTEX_MODE_CDF 3
#define OPTEX_MODE_SORT 4

/* rotation GEMM arithmetic */
#define OPTEX_GEMM_AUTO 0    /* tcgen05 3xTF32 when shapes allow, else fp32 SIMT */
#define OPTEX_GEMM_FP32 1    /* fp32 FFMA tiles (SIMT)                            */
#define OPTEX_GEMM_TF32X3 2  /* tcgen05 kind::tf32, 3-term split, fp32-grade      */
#define OPTEX_GEMM_TF32 3    /* tcgen05 kind::tf32, single pass (reference's own
                                CUDA default, optex.py:248-249)                   */

/* ---- library / device ------------------------------------------------------ */
int optex_abi_version(void);
const char *optex_last_error(void);
/* 0 if the CURRENT cuda device is sm_100 (B200), else OPTEX_EDEVICE. */
int optex_device_check(void);
/* Counts kernels launched by this library since load (bench.py `gpu_launches`). */
uint64_t optex_launch_count(void);
/* Select the rotation-GEMM arithmetic (OPTEX_GEMM_*), process-wide. */
int optex_set_gemm_mode(int gemm_mode);
int optex_get_gemm_mode(void);
/* Programmatic dependent launch between the library's kernels (default on).  Turn it off to time individual
 * kernels with CUDA events: under PDL a kernel may start before its predecessor has drained. */
int optex_set_pdl(int enable);
/* The tensor-core GEMMs split their small operand into tf32 halves in a library-owned scratch buffer: one per device,
 * stream and slot (0..3): work enqueued on different streams never shares scratch.  The slot additionally separates
 * independent pipelines that share one stream; it is a property of the calling host thread (default 0) until changed.
 * optex_ot_step_host_async manages it itself.
 * Returns the previous slot.  (No reference counterpart: the reference runs on one stream, optex.py:256.) */
int optex_set_scratch_slot(int slot);
/* Arithmetic of the Householder construction behind optex_random_rotation(s): 1 = fp64 (default; what the
 * reference's live scipy branch computes in before `.to(pastiche_feature)`, optex.py:147,168), 0 = fp32 (what its
 * impl="torch" branch computes in, optex.py:150-164; 2x faster, |R R^T - I| ~ 1e-6 instead of 1e-8 at c = 512).
 * Returns the previous setting. */
int optex_set_rotation_precision(int fp64);
int optex_get_rotation_precision(void);
/* Debug aid: CTA 0 of every following rotation GEMM writes 64 clock64() stamps (16 k-blocks x {TMA issue, tile
 * landed, hi/lo split done, MMA start}) of its first tile into device_buf (NULL = off). */
int optex_debug_gemm_trace(void *device_buf);

/* ---- the OT step -----------------------------------------------------------
 * replaces: optimal_transport()  optex.py:167-177  (+ the content blend of the
 *           inner loop, optex.py:115-117, as a fused epilogue)
 *
 *   out[n_p, c] = hist_match(P @ R, S @ R, mode) @ R^T      (then, if content:
 *   out += content_strength * (content - out))
 *
 * P [b_p*hw_p, c], S [b_s*hw_s, c], R [c, c] row-major, out [b_p*hw_p, c].
 * b_* / hw_* : batch and pixels-per-sample; the covariance modes take their
 * means per sample (histmatch.py:16,20) and need b_s == 1 or b_s == b_p
 * (histmatch.py:44); cdf/sort pool everything (histmatch.py:11).
 * eps: histmatch.py:5 (the reference always uses 1).
 * content may be NULL.  workspace: optex_ot_workspace_bytes().
 */
size_t optex_ot_workspace_bytes(int64_t n_p, int64_t n_s, int c, int mode);
int optex_ot_step(const float *P, const float *S, const float *R, float *out,
                  int b_p, int64_t hw_p, int b_s, int64_t hw_s, int c, int mode,
                  float eps, const float *content, float content_strength,
                  void *workspace, size_t workspace_bytes, void *stream);

/* Same step with HOST buffers (pageable or pinned): H2D of P, S, R (and
 * content), the step, D2H of out, one stream sync.  Device scratch is owned by
 * the library and grows on demand.  R may be NULL: the rotation is then drawn on
 * the device from (seed, counter) like optex_random_rotation - the reference
 * also draws it inside the call (optex.py:168). */
int optex_ot_step_host(const float *P, const float *S, const float *R, float *out,
                       int b_p, int64_t hw_p, int b_s, int64_t hw_s, int c,
                       int mode, float eps, const float *content,
                       float content_strength, uint64_t seed, uint64_t counter,
                       void *stream);

/* The same without the final synchronisation, for callers that pipeline independent steps: `slot` (0 .. 3)
 * selects one of four library-owned device scratch sets, so step i+1 (slot B, stream B) can upload while step i
 * (slot A, stream A) computes and step i-1 (slot C, stream C) downloads.  Use PINNED host buffers and one stream per
 * slot; the caller synchronises the stream before reading `out` or reusing the slot. */
int optex_ot_step_host_async(const float *P, const float *S, const float *R, float *out,
                             int b_p, int64_t hw_p, int b_s, int64_t hw_s, int c,
                             int mode, float eps, const float *content,
                             float content_strength, uint64_t seed, uint64_t counter,
                             int slot, void *stream);

/* The style block is the same tensor in every iteration of a layer's loop (optex.py:112-113 passes
 * style_features[l] each time): upload it ONCE and keep it resident on the device.  optex_ot_step_host /
 * optex_ot_step_host_async called afterwards with S == NULL (same b_s, hw_s, c) reuse it, on any slot and stream
 * (ordered by an event).  S == NULL here releases the resident block.  One resident block per process. */
int optex_ot_host_set_style(const float *S, int b_s, int64_t hw_s, int c, void *stream);

/* `steps` INDEPENDENT OT steps (optex.py:167-177 each) enqueued by one call - e.g. a batch of syntheses standing at
 * the same layer: step i transports P[(first + i) % n_sets] towards S[(first + i) % n_sets] with rotation i into
 * out[(first + i) % n_out].  P / S / out are HOST arrays of device pointers.
 *   R_all != NULL : rotation i = R_all[i] ([steps, c, c] on the device); R_split (may be NULL) = the tf32 hi / lo planes
 *                   of the same rotations, [steps][2][c][c], from optex_split_rotations - the per-step split launch
 *                   then disappears.  workspace: optex_ot_workspace_bytes().
 *   R_all == NULL : rotation i is drawn on the device from (seed, first_counter + i), like the reference draws one per
 *                   call (optex.py:168) - in batches of up to 32.  workspace: optex_ot_steps_workspace_bytes().
 * Same arithmetic and kernels as `steps` optex_ot_step calls, without a host round trip per step. */
size_t optex_ot_steps_workspace_bytes(int64_t n_p, int64_t n_s, int c, int mode);
int optex_ot_steps(const float *const *P, const float *const *S, int n_sets, const float *R_all,
                   const float *R_split, uint64_t seed, uint64_t first_counter, float *const *out,
                   int n_out, int steps, int first, int b_p, int64_t hw_p, int b_s, int64_t hw_s,
                   int c, int mode, float eps, void *workspace, size_t workspace_bytes, void *stream);
/* out[i] = [tf32_hi(R_i) | tf32_lo(R_i)] for `count` rotations [c, c] (c % 4 == 0) in one launch: the operand halves of
 * the 3xTF32 rotation GEMMs, batched like the draw (optex_random_rotations).  out: [count][2][c][c] floats. */
int optex_split_rotations(const float *R_all, int count, int c, float *out, void *stream);

/* One optex_ot_step with a CUDA event between its stages - the step's own launches, in place, with PDL off for the
 * call; synchronises `stream`.  Measurement aid (no reference counterpart).  Per-channel modes report 5 stages:
 * prepare (split R / reset range slots), forward rotation of P, of S (optex.py:170-171), the matcher
 * (histmatch.py:49-69), inverse rotation (optex.py:175); covariance modes report 1 (the whole step).
 * stage_ms / stage_launches hold OPTEX_MAX_STAGES entries. */
#define OPTEX_MAX_STAGES 8
int optex_ot_step_profile(const float *P, const float *S, const float *R, float *out,
                          int b_p, int64_t hw_p, int b_s, int64_t hw_s, int c, int mode,
                          float eps, void *workspace, size_t workspace_bytes, void *stream,
                          float *stage_ms, int *stage_launches, int *n_stages);

/* An ordinary (not programmatically serialised) empty kernel on `stream`: it starts only once every earlier kernel
 * of the stream has fully drained, so an event recorded behind it brackets PDL-launched work correctly
 * (timing aid; no reference counterpart). */
int optex_fence(void *stream);

/* ---- the inner loop --------------------------------------------------------
 * replaces: optex.py:112-117  (`for _ in range(iters): optimal_transport; blend`)
 * feat [n_p, c] is updated IN PLACE `iters` times.  Rotations: if R_all != NULL
 * it holds iters matrices [iters, c, c]; else they are generated on the device
 * from (seed, first_counter + i)  (optex_random_rotation).
 */
int optex_ot_loop(float *feat, const float *S, const float *R_all, int iters,
                  uint64_t seed, uint64_t first_counter, int b_p, int64_t hw_p,
                  int b_s, int64_t hw_s, int c, int mode, float eps,
                  const float *content, float content_strength, void *workspace,
                  size_t workspace_bytes, void *stream);
size_t optex_ot_loop_workspace_bytes(int64_t n_p, int64_t n_s, int c, int mode);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink ------------------------
 * No reference counterpart (the reference is single-device).  How the step shards:
 *   cdf            PIXEL-sharded - rank g holds n_g rows of P and m_g rows of S (all channels).  The only cross-pixel
 *                  quantities of cdf_match (histmatch.py:52-58) are the per-channel range (all-reduce MIN of 2c words)
 *                  and the two histograms (all-reduce SUM of 2 c 256 counts): every rank builds the tables one GPU
 *                  would, results are BIT-IDENTICAL to the single-GPU step on the concatenated rows, and 1 MB crosses
 *                  NVLink per step at c = 512.
 *   chol/pca/sym   PIXEL-sharded moments (histmatch.py:16-22): column sums and the centred Gram are all-reduced
 *                  (c + c*c floats), the c x c chain runs on every rank, the map is applied to the local rows
 *                  (same result up to fp32 summation order).  One sample per side (b = 1).
 *   sort           needs global ranks per channel: OPTEX_EUNSUPPORTED here (channel-sharded form with an all-gather:
 *                  optimaltextures_b200/parallel.py).
 * NCCL is bound at run time (dlopen of the process's libnccl.so.2).  All ranks must pass the same R (e.g. the same
 * (seed, counter) to optex_random_rotation), c, mode, eps and totals. */
typedef struct optex_comm optex_comm_t;
size_t optex_comm_unique_id_bytes(void);
/* rank 0: create the id (ncclGetUniqueId) and hand it to the other ranks by any means (128 bytes) */
int optex_comm_unique_id(void *id_out, size_t id_bytes);
/* every rank: ncclCommInitRank on the CURRENT device */
int optex_comm_init(const void *id, int rank, int world, optex_comm_t **comm);
/* or wrap a communicator the caller already has (ncclComm_t as void*); it is not destroyed with the handle */
int optex_comm_adopt(void *nccl_comm, int rank, int world, optex_comm_t **comm);
int optex_comm_destroy(optex_comm_t *comm);
int optex_comm_rank(const optex_comm_t *comm);
int optex_comm_world(const optex_comm_t *comm);
/* recv[rank r] = send of rank r (count_per_rank floats each): gathers the local rows of a sharded block */
int optex_comm_allgather_f32(optex_comm_t *comm, const float *send, float *recv, size_t count_per_rank, void *stream);

/* The OT step (optex.py:167-177) on this rank's rows: P_local [n_p_local, c], S_local [n_s_local, c] -> out_local
 * [n_p_local, c]; n_*_total = the sums over the ranks (the cdf histograms and the moments are those of the whole block).
 * content (may be NULL) is this rank's rows of the content block.  workspace: optex_ot_workspace_bytes(n_p_local,
 * n_s_local, c, mode).  out_local must not alias the inputs. */
int optex_ot_step_sharded(optex_comm_t *comm, const float *P_local, const float *S_local, const float *R,
                          float *out_local, int64_t n_p_local, int64_t n_s_local, int64_t n_p_total,
                          int64_t n_s_total, int c, int mode, float eps, const float *content,
                          float content_strength, void *workspace, size_t workspace_bytes, void *stream);

/* ---- histogram matching without rotation -----------------------------------
 * replaces: hist_match()  histmatch.py:5-46  (called directly by
 *           mix_style_features, optex.py:200-201)
 * target [b_t*hw_t, c] and source [b_s*hw_s, c] NHWC-flattened -> out like target.
 */
size_t optex_hist_match_workspace_bytes(int64_t n_t, int64_t n_s, int c, int mode);
int optex_hist_match(const float *target, const float *source, float *out,
                     int b_t, int64_t hw_t, int b_s, int64_t hw_s, int c, int mode,
                     float eps, void *workspace, size_t workspace_bytes,
                     void *stream);

/* ---- per-channel matchers on channel-major data ----------------------------
 * replaces: cdf_match()  histmatch.py:49-69  and  interp()  histmatch.py:72-92
 * target [c, n_t], source [c, n_s] -> out [c, n_t].  Bit-exact with torch-CPU
 * (histc / linspace / cumsum / searchsorted semantics, see oracle/cdf_explicit.py).
 * `tables` (optional, may be NULL): [c, 2, bins] fp32 receiving the upper bin
 * edges and the remapped CDF of every channel (parity hook).  bins <= 1024.
 */
size_t optex_cdf_match_workspace_bytes(int c, int bins);
int optex_cdf_match(const float *target, const float *source, float *out, int c,
                    int64_t n_t, int64_t n_s, int bins, float *tables,
                    void *workspace, size_t workspace_bytes, void *stream);
/* interp(x, xp, fp): x [n], xp/fp [len] -> out [n]  (histmatch.py:72-92). */
int optex_interp(const float *x, const float *xp, const float *fp, float *out,
                 int64_t n, int len, void *stream);

/* Exact 1-D OT per channel (north-star `sort` mode; oracle/sort_oracle.py):
 * stable argsort of target, ascending sort of source, rank r receives
 * sorted_source[((2r+1)*n_s)/(2*n_t)].  `perm` (optional, may be NULL): [c, n_t]
 * int32 receiving the stable sort permutation (parity hook: bit-exact). */
size_t optex_sort_match_workspace_bytes(int c, int64_t n_t, int64_t n_s);
int optex_sort_match(const float *target, const float *source, float *out, int c,
                     int64_t n_t, int64_t n_s, int32_t *perm, void *workspace,
                     size_t workspace_bytes, void *stream);

/* ---- rotations --------------------------------------------------------------
 * replaces: random_rotation()  optex.py:142-164
 * Haar-distributed SO(c) matrix, generated on the device: Philox4x32-10
 * Gaussians keyed by (seed, counter) -> Stewart's Householder construction
 * (the reference's impl="torch" branch, optex.py:151-164) in fp64, stored fp32.
 * If `gauss` != NULL it supplies the normals instead ([c-1, c] fp64, row n uses
 * columns n..c-1) - the parity hook against oracle/rotation.py.
 * R [c, c] row-major.  workspace: optex_rotation_workspace_bytes(c).
 */
size_t optex_rotation_workspace_bytes(int c);
int optex_random_rotation(float *R, int c, uint64_t seed, uint64_t counter,
                          const double *gauss, void *workspace,
                          size_t workspace_bytes, void *stream);

/* `count` rotations R[count][c][c] from counters first_counter .. +count-1 in
 * one batched launch pair (what optex_ot_loop uses: a single matrix is latency
 * bound, a batch fills the machine).  gauss (optional): [count, c-1, c] fp64. */
size_t optex_rotations_workspace_bytes(int c, int count);
int optex_random_rotations(float *R, int c, int count, uint64_t seed,
                           uint64_t first_counter, const double *gauss,
                           void *workspace, size_t workspace_bytes, void *stream);

/* Optional: split R [c,c] into its tf32 hi / lo halves ONCE for the optex_rotate_* calls that follow on this host
 * thread with the same R pointer (3xTF32 arithmetic; optex_ot_step does this internally for its three GEMMs, the
 * three `@` of optex.py:170,171,175 share one R).  The halves live in `workspace` (query the size); the association
 * ends with the next optex_rotation_prepare call - pass R = NULL to release it - and R must not change meanwhile. */
size_t optex_rotation_prepare_workspace_bytes(int c);
int optex_rotation_prepare(const float *R, int c, void *workspace, size_t workspace_bytes, void *stream);

/* ---- rotation GEMMs (building blocks of the step, exported for tests) -------
 * optex.py:170-171:  Xt[c, n] = (X @ R)^T     X [n, c], R [c, c]
 * optex.py:175    :  out[n, c] = Mt^T @ R^T   Mt [c, n] channel-major
 *                    (+ optional content blend, optex.py:117)
 */
int optex_rotate_forward(const float *X, const float *R, float *Xt, int64_t n,
                         int c, void *stream);
/* Channel block of the forward rotation - the unit of the multi-GPU sharding (one block of rotated channels
 * per rank):  Xt[nc, n] = (X @ R[:, c0:c0+nc])^T */
int optex_rotate_forward_block(const float *X, const float *R, float *Xt, int64_t n,
                               int c, int c0, int nc, void *stream);
int optex_rotate_inverse(const float *Mt, const float *R, float *out, int64_t n,
                         int c, const float *content, float content_strength,
                         void *stream);

/* ---- PCA basis of a feature block (the caller either side of the inner loop) -------------------
 * replaces: fit_pca()  optex.py:180-190  and the projections  optex.py:110 (`@ style_eigvs[l]`),
 *           optex.py:120 (`@ style_eigvs[l].T`), optex.py:188 (`tensor @ eigvecs`)
 * X [n, c] NHWC-flattened.  The reference's SVD of X - mean(X) (one scalar mean, optex.py:182) is computed as the
 * eigendecomposition of the c x c Gram matrix, accumulated and diagonalised in FP64 on the device.  Outputs:
 *   eigvecs [c, c] fp32 row-major: column j = right singular vector of the j-th largest singular value; sign
 *                  convention: the component of largest magnitude of every column is positive (an SVD leaves
 *                  the sign open);
 *   sigma   [c]    fp32 singular values, descending (for n < c the trailing c - n are ~0);
 *   k_out   [1]    int32 ON THE DEVICE: the number of leading columns the reference keeps
 *                  (first index where cumsum(sigma / sum(sigma)) > 0.9, optex.py:184).
 * c <= 1024.  Asynchronous like everything else: read k_out after synchronising the stream.
 * Solver (cold solve, basis == NULL): one-sided Jacobi on the rows of G in a BLOCKED order - blocks of 8 rows paired
 * round-robin across CTAs (c / 8 - 1 grid barriers per sweep), all 8 x 8 cross pairs of a block pair rotated inside
 * the CTA between two barriers; no V is carried (the eigenvectors are the normalised rows).  OPTEX_PCA_SOLVER=0
 * selects the round-robin kernel (c - 1 barriers per sweep), which is also what a `basis` request uses.
 * n < c (fewer rows than channels: conv5_1 of a 256^2 pass) is solved in the DUAL form: the n x n matrix A A^T, whose
 * eigenvectors are mapped through A^T; columns n .. c - 1 of eigvecs and sigma[n ..] are 0 (the reference's SVD has
 * only n singular values there, optex.py:183).  OPTEX_PCA_DUAL=0 switches that off. */
size_t optex_fit_pca_workspace_bytes(int64_t n, int c);
int optex_fit_pca(const float *X, int64_t n, int c, float *eigvecs, float *sigma,
                  int32_t *k_out, void *workspace, size_t workspace_bytes, void *stream);
/* The same solve, keeping the eigensolver's FP64 basis between calls (additive; no reference counterpart).
 * optex.py:62-67 refits the PCA of the SAME style image at every pass's resolution; the bases of consecutive passes
 * are close, and a Jacobi iteration started from the previous one converges in about half the sweeps.
 *   basis      [c, c] FP64 on the device, or NULL (then identical to optex_fit_pca).  Rows = the solver's
 *              orthogonal V in its internal (unsorted) order.  Always written with the final V.
 *   warm       0: start from the identity (basis is output only);  != 0: start from the rows of `basis`, which must be
 *              the output of an earlier call with the same c (any orthogonal matrix gives a correct result; a nearby
 *              one gives a fast one).
 *   sweeps_out [1] int32 on the device or NULL: Jacobi sweeps run (observability; 40 = not converged). */
int optex_fit_pca_warm(const float *X, int64_t n, int c, float *eigvecs, float *sigma, int32_t *k_out,
                       double *basis, int warm, int32_t *sweeps_out, void *workspace,
                       size_t workspace_bytes, void *stream);
/* Debug hook (no reference counterpart; scripts/pca_stamps.py): a device buffer of 5 * n int64 that CTA 1 of the
 * blocked-order solver fills with clock64 stamps of its first n global rounds (round start, rows loaded, sub-rounds
 * done, rows stored, barrier passed).  NULL switches it off (the default). */
int optex_debug_pca_stamps(long long *device_buffer, int n);
/* The same for the cooperative covariance-chain kernel (cov_chain.cu; scripts/chain_stamps.py): n int64 stamps of
 * CTA 0, one at kernel entry, one behind every grid barrier, one at the end.  NULL switches it off (the default). */
int optex_debug_chain_stamps(long long *device_buffer, int n);
/* transpose = 0: out[n, k] = X[n, c] V[c, k]      (project onto the basis,  optex.py:110, :188)
 * transpose = 1: out[n, c] = X[n, k] V[c, k]^T    (back to feature space,   optex.py:120)
 * V dense row-major [c, k].  Tensor cores when c and k allow (k % 32 == 0, c % 4 == 0), else fp32 SIMT tiles. */
int optex_pca_project(const float *X, const float *V, float *out, int64_t n, int c, int k,
                      int transpose, void *stream);

/* ---- VGG-19 encoder / feature-inverter layers ----------------------------------------------------
 * replaces: the nn.Sequential stacks of vgg.py:14-136 behind Encoder.forward (vgg.py:152-153, called at
 *           optex.py:62-63,107) and Decoder.forward (vgg.py:170-171, called at optex.py:122), one layer per call:
 *   dst[b, h, w, c_out] = [relu]( conv3x3( reflection_pad1( pre_op(src) ) ) + bias )
 * pre_op: OPTEX_PRE_NONE | OPTEX_PRE_MAXPOOL2 (MaxPool2d(2, 2, ceil_mode=True): h = ceil(h_src / 2)) |
 *         OPTEX_PRE_UPSAMPLE2 (UpsamplingNearest2d(2): h = 2 h_src); neither is materialised.
 * src: NHWC [b, h_src, w_src, c_in] (or the NCHW image when src_nchw != 0).  dst: NHWC with row pitch ldd >= c_out.
 * weight: [c_out, kp] row-major, kp = optex_conv3x3_packed_k(c_in) = 9 c_in rounded up to 32, element
 *         [co, (ky * 3 + kx) * c_in + ci] = torch weight [co, ci, ky, kx], the padding columns zero.
 * bias: [c_out] or NULL.  Arithmetic: the library's GEMM mode (3xTF32 by default).  h, w >= 2 after the pre-op. */
#define OPTEX_PRE_NONE 0
#define OPTEX_PRE_MAXPOOL2 1
#define OPTEX_PRE_UPSAMPLE2 2
int optex_conv3x3_packed_k(int c_in);
size_t optex_conv3x3_workspace_bytes(int b, int h_src, int w_src, int c_in, int c_out, int pre_op);
int optex_conv3x3(const float *src, int src_nchw, int b, int h_src, int w_src, int c_in,
                  const float *weight, const float *bias, int c_out, int pre_op, int relu,
                  float *dst, int64_t ldd, void *workspace, size_t workspace_bytes, void *stream);
/* dst[b, c, hw] = src[b, hw, 0:c], src row pitch c_src >= c  (the decoder's image output, vgg.py:171 is NCHW) */
int optex_nhwc_to_nchw(const float *src, float *dst, int b, int64_t hw, int c_src, int c, void *stream);

/* ---- image-space glue of one synthesis pass -------------------------------------------------
 * replaces: resize()  util.py:105-106 = interpolate(x, size, mode="bicubic", align_corners=False, antialias=True),
 *           called at optex.py:48,51,56 (multires re-scaling of styles / content / pastiche between passes).
 * src [planes, h_in, w_in] -> dst [planes, h_out, w_out], planes = b * c of an NCHW tensor.  Separable, width
 * first; window and weights as torch computes them (cubic a = -0.5, support widened by the scale when shrinking). */
size_t optex_resize_workspace_bytes(int planes, int h_in, int w_in, int h_out, int w_out);
int optex_resize_bicubic_aa(const float *src, float *dst, int planes, int h_in, int w_in, int h_out,
                            int w_out, void *workspace, size_t workspace_bytes, void *stream);
/* replaces: kornia.color.hls.rgb_to_hls / hls_to_rgb as used by the colour transfer, optex.py:126-128.
 * NCHW [b, 3, hw] fp32; H in radians [0, 2 pi), L and S in [0, 1]; undefined hue / saturation (grey) -> 0.
 * kornia is not vendored by the reference and absent from the build image: parity UNPINNED (oracle restates the
 * published formulas).  optex_lightness_transfer = hls_to_rgb(H(content), L(pastiche), S(content)) in one pass. */
int optex_rgb_to_hls(const float *rgb, float *hls, int b, int64_t hw, void *stream);
int optex_hls_to_rgb(const float *hls, float *rgb, int b, int64_t hw, void *stream);
int optex_lightness_transfer(const float *content, const float *pastiche, float *out, int b, int64_t hw,
                             void *stream);
/* replaces: the per-layer body of mix_style_features()  optex.py:197-204 (the two hist_match calls are
 * optex_hist_match):  out = (A (1-a) + AtoB a) m + (BtoA (1-a) + B a) (1-m),  m = mask [mask_h, mask_w] resized
 * to [h, w] with interpolate(mode="nearest").  A, B, AtoB, BtoA, out: [h, w, c]. */
int optex_mix_features(const float *A, const float *B, const float *AtoB, const float *BtoA,
                       const float *mask, float *out, int h, int w, int c, int mask_h, int mask_w,
                       float alpha, void *stream);
/* replaces: optex.py:76  content_feature - content_feature.mean() + torch.mean(style_features[l])
 * (scalar means over the whole tensors, accumulated in FP64, deterministic). */
size_t optex_recentre_workspace_bytes(void);
int optex_recentre(const float *content, int64_t n_content, const float *style, int64_t n_style,
                   float *out, void *workspace, size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* OPTEX_B200_H */
