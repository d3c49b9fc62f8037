/*
 * spmm_oracle.c — CPU restatement of the reference's SpMM hot path.  TEST INFRASTRUCTURE ONLY:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this; the product (libspmm_b200.so, custom_mm, matmuls) never does.
 *
 * Parity status: PINNED — against outputs of the real reference recorded on a B200
 * (tests/golden/ref_*.npz, tests/golden/make_ref_golden.py: cuSPARSE forward, naive kernel, grad_m2, CSR arrays)
 * and against fixtures generated from the reference's own Python layer in the build container
 * (tests/golden/make_golden.py imports /root/reference/matmuls.py and records what
 * `a.to_sparse_csr()` + its int32 casts and `sparse_matmul` produce) — see
 * tests/test_oracle.py.  The reference's arithmetic itself lives in cuSPARSE 12.5
 * (closed source, cusparseSpMM CSR_ALG2, src/baseline_mm.cu:207); its only first-party
 * statement is the naive kernel, which this file follows line by line.
 *
 * Citations are relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* C[r,:] = sum_{p in [row_ptr[r], row_ptr[r+1])} val[p] * B[col[p], :], empty row -> 0.
 * src/naive_sparse_mm.cu:52-99 (row loop :60-92, update = SUM, write :95-99), evaluated with
 * an fp64 accumulator: this is the "fp64 oracle" of BASELINE.json's tolerance. */
void oracle_spmm_f64(const int32_t* row_ptr, const int32_t* col, const float* val, int64_t M,
                     int64_t N, const float* B, int64_t ldb, double* C /* [M,N] */) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t r = 0; r < M; ++r) {
    double* c = C + r * N;
    for (int64_t j = 0; j < N; ++j) c[j] = 0.0;
    for (int32_t p = row_ptr[r]; p < row_ptr[r + 1]; ++p) {
      const double v = (double)val[p];
      const float* b = B + (int64_t)col[p] * ldb;
      for (int64_t j = 0; j < N; ++j) c[j] += v * (double)b[j];
    }
  }
}

/* Same loop nest in fp32, products accumulated in CSR order exactly like the naive kernel
 * (src/naive_sparse_mm.cu:84-90: val = vals[i] * mat[...]; result += val).  `fused` selects
 * fmaf (what nvcc contracts the two statements into) or separate multiply and add. */
void oracle_spmm_f32(const int32_t* row_ptr, const int32_t* col, const float* val, int64_t M,
                     int64_t N, const float* B, int64_t ldb, float* C, int fused) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t r = 0; r < M; ++r) {
    float* c = C + r * N;
    for (int64_t j = 0; j < N; ++j) c[j] = 0.0f;
    for (int32_t p = row_ptr[r]; p < row_ptr[r + 1]; ++p) {
      const float v = val[p];
      const float* b = B + (int64_t)col[p] * ldb;
      if (fused) {
        for (int64_t j = 0; j < N; ++j) c[j] = fmaf(v, b[j], c[j]);
      } else {
        for (int64_t j = 0; j < N; ++j) {
          volatile float t = v * b[j];
          c[j] = c[j] + t;
        }
      }
    }
  }
}

/* Dense row-major [M,K] -> CSR.  Predicate x != 0 (NaN kept, -0.0 dropped), row-major order,
 * int32 indices: what `a.to_sparse_csr()` (matmuls.py:215,228) followed by the int32 casts of
 * get_sparse_tensor_properties (matmuls.py:185-186) yields, and what cusparseDenseToSparse
 * (src/baseline_mm.cu:241-258) is documented to produce.  Two calls like the cuSPARSE pair:
 * count fills row_ptr and returns nnz; fill writes col / val. */
int64_t oracle_dense_to_csr_count(const float* dense, int64_t M, int64_t K, int64_t ld,
                                  int32_t* row_ptr) {
  int64_t n = 0;
  row_ptr[0] = 0;
  for (int64_t r = 0; r < M; ++r) {
    const float* x = dense + r * ld;
    for (int64_t c = 0; c < K; ++c) n += (x[c] != 0.0f);
    row_ptr[r + 1] = (int32_t)n;
  }
  return n;
}

void oracle_dense_to_csr_fill(const float* dense, int64_t M, int64_t K, int64_t ld, int32_t* col,
                              float* val) {
  int64_t n = 0;
  for (int64_t r = 0; r < M; ++r) {
    const float* x = dense + r * ld;
    for (int64_t c = 0; c < K; ++c) {
      if (x[c] != 0.0f) {
        col[n] = (int32_t)c;
        memcpy(&val[n], &x[c], sizeof(float)); /* bit copy: NaN payloads survive */
        ++n;
      }
    }
  }
}

/* CSR(A) -> CSR(A^T) with ascending original row ids inside every new row — the canonical
 * CSR that `m1.transpose(-1, -2).to_sparse_csr()` builds in backward (matmuls.py:253). */
void oracle_csr_transpose(const int32_t* row_ptr, const int32_t* col, const float* val, int64_t M,
                          int64_t K, int32_t* t_row_ptr, int32_t* t_col, float* t_val) {
  const int64_t nnz = row_ptr[M];
  for (int64_t c = 0; c <= K; ++c) t_row_ptr[c] = 0;
  for (int64_t p = 0; p < nnz; ++p) t_row_ptr[col[p] + 1]++;
  for (int64_t c = 0; c < K; ++c) t_row_ptr[c + 1] += t_row_ptr[c];
  int32_t* cursor = (int32_t*)malloc(sizeof(int32_t) * (size_t)(K > 0 ? K : 1));
  for (int64_t c = 0; c < K; ++c) cursor[c] = t_row_ptr[c];
  for (int64_t r = 0; r < M; ++r) {
    for (int32_t p = row_ptr[r]; p < row_ptr[r + 1]; ++p) {
      const int32_t q = cursor[col[p]]++;
      t_col[q] = (int32_t)r;
      memcpy(&t_val[q], &val[p], sizeof(float));
    }
  }
  free(cursor);
}

/* grad_A on A's pattern: out[p] = sum_j G[row(p), j] * B[col(p), j] in fp64 — the formula
 * grad_m1 = grad_output x m2^T of matmuls.py:250, sampled at A's non-zeros. */
void oracle_sddmm_f64(const int32_t* row_ptr, const int32_t* col, int64_t M, int64_t N,
                      const float* G, int64_t ldg, const float* B, int64_t ldb, double* out) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t r = 0; r < M; ++r) {
    const float* g = G + r * ldg;
    for (int32_t p = row_ptr[r]; p < row_ptr[r + 1]; ++p) {
      const float* b = B + (int64_t)col[p] * ldb;
      double s = 0.0;
      for (int64_t j = 0; j < N; ++j) s += (double)g[j] * (double)b[j];
      out[p] = s;
    }
  }
}

/* grad_B = A^T * G (matmuls.py:253) in fp64, scatter form over A's CSR. Serial (scatter). */
void oracle_spmm_t_f64(const int32_t* row_ptr, const int32_t* col, const float* val, int64_t M,
                       int64_t K, int64_t N, const float* G, int64_t ldg, double* dB /* [K,N] */) {
  for (int64_t i = 0; i < K * N; ++i) dB[i] = 0.0;
  for (int64_t r = 0; r < M; ++r) {
    const float* g = G + r * ldg;
    for (int32_t p = row_ptr[r]; p < row_ptr[r + 1]; ++p) {
      double* d = dB + (int64_t)col[p] * N;
      const double v = (double)val[p];
      for (int64_t j = 0; j < N; ++j) d[j] += v * (double)g[j];
    }
  }
}
