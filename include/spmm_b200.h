/*
 * spmm_b200.h — C ABI of the B200-native fp32 CSR SpMM engine (libspmm_b200.so).
 *
 * This is the drop-in boundary for the reference's SpMM hot path.  Every entry
 * point takes plain pointers and sizes (no torch types); device pointers are
 * what the reference's pybind layer extracts with data_ptr<T>() before it calls
 * its own wrappers.  Citations are relative to the reference tree
 * (smoorjani/matrix-multiplication):
 *
 *   b200_spmm_csr_forward      replaces cusparse_mm_wrapper      src/baseline_mm.cu:167-216
 *                              and      naive_spmm_wrapper       src/naive_sparse_mm.cu:104-136
 *                              (bound as custom_mm.cusparse_mmul src/custom_mm.cpp:203-217,
 *                               custom_mm.naive_spmm             src/custom_mm.cpp:166-179)
 *   b200_dense_to_csr_*        replaces dense_to_csr / free_csr  src/baseline_mm.cu:218-270
 *   b200_csr_transpose         replaces m1.transpose().to_sparse_csr()   matmuls.py:253
 *   b200_sddmm_csr             replaces the dense grad_m1 product        matmuls.py:250
 *   b200_plan_*                replaces the layer-keyed inspect/execute/clean registries
 *                              src/custom_mm.cpp:236-281 (cusparse_*) and :293-357 (tiledspmm_*)
 *
 * Conventions (same as the reference, src/custom_mm.cpp:203-217):
 *   - fp32 values, int32 indices, index base 0, CSR with sorted rows;
 *   - dense matrices are ROW-MAJOR with an explicit leading dimension in elements;
 *   - C is fully overwritten (alpha = 1, beta = 0), empty rows produce zeros;
 *   - the caller owns every buffer; the library only borrows pointers for the call
 *     (a plan borrows the CSR arrays until b200_plan_destroy);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     device entry points never synchronise the stream unless documented.
 *
 * Error behaviour: every function returns 0 on success and a non-zero
 * b200_status otherwise; b200_last_error() returns a thread-local message.
 * (The reference throws std::logic_error on CUDA failures, src/Utilities.cuh:17-22,
 *  and silently prints cuSPARSE failures, :145-154; here every failure is reported.)
 */
#ifndef SPMM_B200_H_
#define SPMM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  B200_OK = 0,
  B200_ERR_INVALID = 1,   /* bad argument (null pointer, negative size, ld < N, ...) */
  B200_ERR_CUDA = 2,      /* a CUDA runtime call or kernel launch failed            */
  B200_ERR_NO_DEVICE = 3, /* no sm_100 device available                             */
  B200_ERR_ALLOC = 4      /* device or host allocation failed                       */
} b200_status;

/* Thread-local description of the last failure on this thread ("" if none). */
const char* b200_last_error(void);

/* Library/ABI version (bumped when a signature changes) and build info string. */
int b200_abi_version(void);
const char* b200_build_info(void);

/* Number of SpMM-engine kernels launched by this process so far (all threads). */
uint64_t b200_kernel_launch_count(void);

/* Launches one empty kernel on `stream`, waits for it and returns the SM count:
 * "is there a usable sm_100 device" (stands in for the reference's dummy_kernel,
 * src/baseline_mm.cu:24-35, and for cusparseCreate in init_cusparse). */
int b200_smoke_kernel(void* stream, int* sm_count);

/* ---- tuning knobs (process-wide; for benchmarks and tests) ------------------
 * stage_mode: 0 = col/val read straight from global memory,
 *             1 = col/val segments staged in shared memory with TMA bulk copies
 *                 (cp.async.bulk + mbarrier); default 1.
 * items_per_task: merge-path items (row ends + non-zeros) per warp task; 0 = auto (256, less for small problems).
 * pipeline      : forward kernel variant, 0 = shipped (profiles/r1_variant_sweep.md lists the alternatives).
 * flat_min_n    : smallest N served by the one-lane-group kernel (default 24; below it lane groups share a warp).
 * slab_mode     : narrow-N (4 / 8 / 16 columns) shared-memory slab kernel: 0 never, 1 auto (default).
 * tc_mode       : block-dense tensor-core path: 0 never, 1 auto (default), 2 whenever the shape allows it.
 * tc_min_density_pct : auto rule floor, nnz / (M*K) in percent (default 2); above it a measured cost model of the two
 *                 paths decides (small or very sparse problems stay on the gather kernels).
 * tc_kind       : split-precision element kind: 0 = tf32 hi + lo, 1 = fp16 hi + lo with per-row / per-column scaling.
 * tc_nt, tc_terms, tc_chunk : tile width override (multiple of 64), which split-precision products are issued (bit mask,
 *                 7 = all), k-tiles accumulated in TMEM between two fp32 register accumulations (accuracy studies).
 * tc_pair       : CTA pairs (tcgen05 cta_group::2) for the tiles: 0 never, 1 whenever the shape allows it, 2 auto (default:
 *                 256-wide tiles on a long contraction).
 * tc_sddmm_short: tensor-core SDDMM whose contraction fits one TMEM accumulation (N <= tc_chunk k-tiles): 1 = the
 *                 transposed-product kernel, tile shape chosen by contraction depth and pattern density (default),
 *                 2 / 3 = that kernel with 128- / 256-row tiles, 0 = the general kernel. */
int b200_set_option(const char* name, int64_t value);
int64_t b200_get_option(const char* name);

/* ---- workspace sizing --------------------------------------------------------
 * Bytes of device scratch b200_spmm_csr_forward needs for a matrix with M rows,
 * nnz non-zeros and N dense columns (partial sums of rows split across tasks). */
size_t b200_spmm_workspace_bytes(int64_t M, int64_t nnz, int64_t N);

/* ---- forward: C[M,N] = A_csr[M,K] * B[K,N] ------------------------------------
 * All pointers are DEVICE pointers.  workspace may be NULL (then the library
 * uses a cached per-device scratch buffer grown on demand, stream-ordered). */
int b200_spmm_csr_forward(const int32_t* d_row_ptr, const int32_t* d_col_ind,
                          const float* d_val, int64_t M, int64_t K, int64_t nnz,
                          const float* d_B, int64_t ldb, int64_t N,
                          float* d_C, int64_t ldc,
                          void* d_workspace, size_t workspace_bytes, void* stream);

/* Same, with per-call flags (B200_FLAG_NO_STAGE: read col/val straight from global
 * memory instead of staging them in shared memory with TMA bulk copies). */
#define B200_FLAG_NO_STAGE 1u
int b200_spmm_csr_forward_flags(const int32_t* d_row_ptr, const int32_t* d_col_ind,
                                const float* d_val, int64_t M, int64_t K, int64_t nnz,
                                const float* d_B, int64_t ldb, int64_t N,
                                float* d_C, int64_t ldc,
                                void* d_workspace, size_t workspace_bytes, void* stream,
                                uint32_t flags);

/* ---- SDDMM on A's pattern: out_val[p] = sum_j G[row(p), j] * B[col(p), j] ------
 * grad_A of C = A*B restricted to A's non-zeros (matmuls.py:250, intended formula
 * grad_C * B^T sampled at the pattern).  G is [M,N] (ldg), B is [K,N] (ldb). */
int b200_sddmm_csr(const int32_t* d_row_ptr, const int32_t* d_col_ind,
                   int64_t M, int64_t K, int64_t nnz,
                   const float* d_G, int64_t ldg, const float* d_B, int64_t ldb, int64_t N,
                   float* d_out_val, void* stream);

/* ---- CSR transpose (CSR of A^T, i.e. CSC of A), rows sorted, stable ------------
 * Outputs: t_row_ptr[K+1], t_col_ind[nnz] (original row ids), t_val[nnz], and
 * optionally t_perm[nnz] (t_val[q] = val[t_perm[q]]; may be NULL).
 * d_val may be NULL (pattern only; then t_val is ignored).
 * Synchronises `stream` internally only for temporary allocation bookkeeping. */
int b200_csr_transpose(const int32_t* d_row_ptr, const int32_t* d_col_ind, const float* d_val,
                       int64_t M, int64_t K, int64_t nnz,
                       int32_t* d_t_row_ptr, int32_t* d_t_col_ind, float* d_t_val,
                       int32_t* d_t_perm, void* stream);

/* Re-gather values through a transpose permutation: dst[q] = src[perm[q]]. */
int b200_gather_f32(const float* d_src, const int32_t* d_perm, int64_t n, float* d_dst, void* stream);

/* ---- dense -> CSR, bit-exact with the reference's construction ----------------
 * Keeps x != 0.0f (NaN kept, -0.0 dropped), row-major order, int32 indices.
 * Two-phase like cusparseDenseToSparse_analysis/_convert (src/baseline_mm.cu:241-258):
 *   count: writes d_row_ptr[M+1] (exclusive scan, row_ptr[M] = nnz) and, if
 *          h_nnz != NULL, synchronises the stream and returns nnz on the host;
 *   fill : writes d_col_ind[nnz], d_val[nnz] for a row_ptr produced by count. */
int b200_dense_to_csr_count(const float* d_dense, int64_t M, int64_t K, int64_t ld,
                            int32_t* d_row_ptr, int64_t* h_nnz, void* stream);
int b200_dense_to_csr_fill(const float* d_dense, int64_t M, int64_t K, int64_t ld,
                           const int32_t* d_row_ptr, int32_t* d_col_ind, float* d_val,
                           void* stream);
/* One pass: count, global prefix (decoupled look-back across CTAs) and fill in ONE kernel — the dense matrix is read
 * exactly once (the two-phase pair above reads it twice, as cusparseDenseToSparse_analysis + _convert do,
 * src/baseline_mm.cu:241-258).  The caller provides room for `capacity` non-zeros (M*K always suffices; a known
 * sparsity gives a tighter bound).  Writes row_ptr[M+1] and the first min(nnz, capacity) entries of col_ind / val.
 * If h_nnz != NULL the stream is synchronised and nnz returned; nnz > capacity is reported as B200_ERR_INVALID with
 * *h_nnz set and row_ptr complete, so the caller can allocate exactly and finish with b200_dense_to_csr_fill.
 * d_scratch: b200_dense_to_csr_onepass_scratch_bytes(M, K) bytes of device memory, 16-byte aligned (tile states of
 * the look-back); NULL = the library allocates and frees it around the call (and synchronises the stream).
 * Same bit-exact contract as above. */
size_t b200_dense_to_csr_onepass_scratch_bytes(int64_t M, int64_t K);
int b200_dense_to_csr_onepass(const float* d_dense, int64_t M, int64_t K, int64_t ld, int32_t* d_row_ptr,
                              int32_t* d_col_ind, float* d_val, int64_t capacity, void* d_scratch,
                              size_t scratch_bytes, int64_t* h_nnz, void* stream);
/* One-call variant with the reference's ownership (dense_to_csr allocates with
 * cudaMalloc, free_csr releases; src/baseline_mm.cu:218-270). Synchronises. */
int b200_dense_to_csr_alloc(const float* d_dense, int64_t M, int64_t K, int64_t ld,
                            float** d_val, int32_t** d_col_ind, int32_t** d_row_ptr,
                            int64_t* nnz, void* stream);
int b200_free_csr(float* d_val, int32_t* d_col_ind, int32_t* d_row_ptr);

/* Batched operands: a stack of L dense [M,K] matrices converted as ONE [L*M, K] matrix (b200_dense_to_csr_*)
 * becomes the block-diagonal CSR of shape [L*M, L*K] by shifting the columns of row r by (r / M) * K.
 * One SpMM against the stacked dense operand [L*K, N] then replaces the per-sample loop of the reference
 * (matmuls.py:221-225, 289-293).  In place on d_col_ind; L*K must fit int32. */
int b200_csr_block_diag_cols(const int32_t* d_row_ptr, int32_t* d_col_ind, int64_t rows, int64_t M, int64_t K,
                             void* stream);

/* ---- plans: the inspect / execute / clean triple -------------------------------
 * A plan borrows a device CSR, owns its scratch, and lazily builds + caches the
 * transposed CSR used by backward_B.  Not thread-safe per plan.  A plan belongs to the device that was current
 * when it was created: calls made with another current device return B200_ERR_INVALID. */
typedef struct b200_plan b200_plan;

int b200_plan_create(const int32_t* d_row_ptr, const int32_t* d_col_ind, const float* d_val,
                     int64_t M, int64_t K, int64_t nnz, b200_plan** out_plan, void* stream);
int b200_plan_destroy(b200_plan* plan);
/* Tell the plan that d_val changed in place (pattern unchanged): the cached
 * transposed values are re-gathered on next use. */
int b200_plan_values_updated(b200_plan* plan);
/* max row length, number of empty rows (computed at creation). */
int b200_plan_stats(const b200_plan* plan, int64_t* max_row_nnz, int64_t* empty_rows);

int b200_plan_forward(b200_plan* plan, const float* d_B, int64_t ldb, int64_t N,
                      float* d_C, int64_t ldc, void* stream);
/* Multi-GPU forward (row-sharded A, one plan per rank; the reference is single-GPU, src/sparse_mm.cu:295):
 * same as b200_plan_forward, and every finished row of C is ALSO stored into the same row block of the other
 * ranks' buffers from the kernel's epilogue, so the all-gather of C overlaps the product instead of following it.
 *   peer_C   : n_peers (<= 7) device pointers, mapped into this process (CUDA IPC / symmetric memory), each
 *              pointing at THIS rank's row block inside a peer's C buffer (same ldc); may be NULL if n_peers == 0;
 *   mc_C     : optional NVLS multicast address of this rank's row block (multimem.st: the NVSwitch replicates the
 *              store to every rank, this one included); when non-NULL it replaces d_C and peer_C as the target.
 * The caller synchronises the ranks afterwards (a barrier) before anyone reads a peer-written row.  N >= 24. */
int b200_plan_forward_peers(b200_plan* plan, const float* d_B, int64_t ldb, int64_t N,
                            float* d_C, int64_t ldc, float* const* peer_C, int n_peers, float* mc_C,
                            void* stream);
/* ---- extended forward: block-dense tensor-core tiles, fused epilogue, transposed operands ---------------------
 * Same product as b200_plan_forward with the options a sparse FC layer needs (the caller being replaced is
 * benchmarks/cusparse_fc_layer.py:41-46: `cusparseMM.apply(inp, weight.t()) + bias`, which pays two transposes
 * and a separate bias add per call):
 *   B200_OP_B_TRANSPOSED : d_B is B^T, row-major [N, K] with leading dimension ldb (activations [tokens, features]);
 *   B200_OP_C_TRANSPOSED : the result is written as C^T, row-major [N, M] with leading dimension ldc;
 *   B200_OP_RELU         : max(x, 0) applied in the epilogue;
 *   d_bias               : optional [M] vector added to every column of C (NULL = none);
 *   B200_OP_NO_TC / B200_OP_FORCE_TC : override the path choice below.
 * Path choice (also applied by b200_plan_forward / b200_plan_backward_B): when a measured cost model says the
 * tiles win (large enough problem, density above ~4-5 %, N >= 8), the plan densifies the non-empty 128 x 32 tiles of
 * A once (fp32 split into tf32 hi + lo) and the product runs on the tensor cores (tcgen05.mma kind::tf32, three
 * products per k-step, fp32 accumulation in TMEM): fp32-grade accuracy (rel-err <= 1e-5 vs fp64 is tested), 2-6x the
 * gather kernels' throughput.  This replaces the reference's cuSPARSE call for dense-ish operands
 * (src/baseline_mm.cu:186-213) and is what its dense comparison path (src/baseline_mm.cu:52-102) is measured against.
 * Inf/NaN in the dense operand are detected on the device and that call runs on the CSR order instead (a dense
 * contraction would turn 0 * Inf into NaN).  Otherwise the gather kernels run. */
#define B200_OP_B_TRANSPOSED 1u
#define B200_OP_C_TRANSPOSED 2u
#define B200_OP_RELU 4u
#define B200_OP_NO_TC 8u
#define B200_OP_FORCE_TC 16u
int b200_plan_forward_ex(b200_plan* plan, const float* d_B, int64_t ldb, int64_t N, float* d_C, int64_t ldc,
                         const float* d_bias, uint32_t flags, void* stream);
/* grad_B = A^T * G with the same operand options (d_G may be G^T [N, M], d_dB may be written as dB^T [N, K]). */
int b200_plan_backward_B_ex(b200_plan* plan, const float* d_G, int64_t ldg, int64_t N, float* d_dB, int64_t lddb,
                            uint32_t flags, void* stream);
/* grad of the CSR values with the same path choice: on dense-enough operands D = G * B^T runs on the tensor-core tiles
 * (both operands packed per call, contraction over the N dense columns) and the epilogue writes only the entries on
 * A's pattern.  B200_OP_B_TRANSPOSED: BOTH operands are given transposed, d_G as [N, M] and d_B as [N, K] (the
 * gradient and the activations of an FC layer as they are; served by the tiles, or by a plain CSR-order kernel that is
 * only meant for small problems). */
int b200_plan_backward_A_ex(b200_plan* plan, const float* d_G, int64_t ldg, const float* d_B, int64_t ldb, int64_t N,
                            float* d_dval, uint32_t flags, void* stream);
/* 1 if a product with N dense columns over this side of the plan (0 = A: forward, 1 = A^T: backward_B, 2 = the SDDMM
 * of backward_A) would take the tensor-core tiles
 * under the current options and flags, else 0.  Callers that hold their operands transposed use it to decide whether
 * to pass them as they are (B200_OP_*_TRANSPOSED are served efficiently by the tiles only). */
int b200_plan_would_use_tc(const b200_plan* plan, int side, int64_t N, uint32_t flags);
/* Block-dense state of a plan side (0 = A, 1 = A^T): stored tiles, their bytes, and the path the last product took
 * (0 gather kernels, 1 tensor-core tiles, 2 strided fallback, 3 shared-memory slab kernel for narrow N). */
int b200_plan_tc_info(const b200_plan* plan, int side, int64_t* n_tiles, int64_t* tile_bytes, int* last_path);
/* Debug aid: non-zero when a pipeline wait of the last tensor-core launch gave up (protocol bug); synchronises. */
int b200_tc_last_timeout(const b200_plan* plan, int* code);

/* grad_B[K,N] = A^T * G[M,N]  — SpMM over the cached transposed CSR (matmuls.py:253). */
int b200_plan_backward_B(b200_plan* plan, const float* d_G, int64_t ldg, int64_t N,
                         float* d_dB, int64_t lddb, void* stream);
/* grad_A values on A's pattern (SDDMM). */
int b200_plan_backward_A(b200_plan* plan, const float* d_G, int64_t ldg,
                         const float* d_B, int64_t ldb, int64_t N,
                         float* d_dval, void* stream);

/* End-to-end forward with HOST buffers on a resident plan: C_host[M,N] = A * B_host[K,N].
 * B is cut into `panels` column panels (0 = auto); the upload of panel p+1, the product of panel p and the
 * download of panel p-1 run concurrently on three streams (PCIe is full duplex; C[:, panel] needs only
 * B[:, panel]), so a step costs about max(H2D, D2H) instead of H2D + product + D2H.  h_B / h_C should be pinned
 * (cudaHostAlloc / cudaHostRegister) — pageable memory works but serialises.  Returns after C is in host memory. */
int b200_plan_forward_host(b200_plan* plan, const float* h_B, int64_t N, float* h_C, int panels);

/* ---- host-buffer convenience entry points (end-to-end path) --------------------
 * All pointers are HOST pointers (pinned memory makes the copies asynchronous).
 * Each call copies its inputs to the device, runs the device path above and
 * copies the result back; it returns after the result is in host memory. */
int b200_spmm_csr_forward_host(const int32_t* h_row_ptr, const int32_t* h_col_ind,
                               const float* h_val, int64_t M, int64_t K, int64_t nnz,
                               const float* h_B, int64_t N, float* h_C);
int b200_sddmm_csr_host(const int32_t* h_row_ptr, const int32_t* h_col_ind,
                        int64_t M, int64_t K, int64_t nnz,
                        const float* h_G, const float* h_B, int64_t N, float* h_out_val);
int b200_spmm_csr_backward_B_host(const int32_t* h_row_ptr, const int32_t* h_col_ind,
                                  const float* h_val, int64_t M, int64_t K, int64_t nnz,
                                  const float* h_G, int64_t N, float* h_dB);
int b200_csr_transpose_host(const int32_t* h_row_ptr, const int32_t* h_col_ind, const float* h_val,
                            int64_t M, int64_t K, int64_t nnz,
                            int32_t* h_t_row_ptr, int32_t* h_t_col_ind, float* h_t_val);
/* Returns nnz through *nnz; h_col_ind/h_val must hold capacity elements
 * (B200_ERR_INVALID with *nnz set if capacity is too small; pass capacity 0 to query). */
int b200_dense_to_csr_host(const float* h_dense, int64_t M, int64_t K,
                           int32_t* h_row_ptr, int32_t* h_col_ind, float* h_val,
                           int64_t capacity, int64_t* nnz);

/* ---- host-side partition helpers (pure CPU; used by planners and tests) --------
 * Merge-path task boundary for diagonal d over (row ends, non-zeros):
 * returns the (row, nnz offset) a task starting at item d begins with, snapped to
 * the row start unless the row is longer than heavy_threshold. */
int b200_merge_path_boundary(const int32_t* h_row_ptr, int64_t M, int64_t nnz,
                             int64_t d, int64_t heavy_threshold,
                             int64_t* row, int64_t* nz);
/* nnz-balanced contiguous row shards: shard g owns rows [bounds[g], bounds[g+1]).
 * cost(row range) = nnz * nnz_weight + rows * row_weight (row_weight ~ N for C writes). */
int b200_shard_rows(const int32_t* h_row_ptr, int64_t M, int parts,
                    double nnz_weight, double row_weight, int64_t* bounds /* parts+1 */);

#ifdef __cplusplus
}
#endif
#endif /* SPMM_B200_H_ */
