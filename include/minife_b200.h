/* include/minife_b200.h -- C ABI of libminife_b200.so, the B200-native (sm_100a) implementation of the
 * miniFE hot path: hex8 FE assembly + unpreconditioned CG.  Plain pointers and sizes only.
 *
 * Each entry point names the reference interface it replaces (paths relative to the Mantevo/miniFE tree).
 * One mfe_ctx == one MPI rank of the reference == one GPU.  Device state (CSR arrays, vectors, halo
 * buffers) lives in the context; host arrays are only touched by the *_export / *_import calls and by
 * the host-buffer arguments documented below.  All functions return 0 on success, nonzero on error
 * (message via mfe_last_error()).  Every phase call returns after its stream is synchronised, so host
 * wall clocks around a call (ref/src/driver.hpp:74-88 RUN_TIMED_FUNCTION) are honest.
 *
 * There is no CPU fallback: without a CUDA device every compute entry point fails with an error.
 */
#ifndef MINIFE_B200_H
#define MINIFE_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mfe_ctx mfe_ctx;

#define MFE_ABI_VERSION 1

/* cg_times[] slots, same order as ref/utils/mytimer.hpp:37-44 */
enum { MFE_WAXPY = 0, MFE_DOT = 1, MFE_MATVEC = 2, MFE_MATVECDOT = 3, MFE_TOTAL = 4, MFE_NUM_TIMERS = 5 };

/* dot-product algorithm (ref/src/Vector_functions.hpp:194-226 sums strictly left to right):
 *   MFE_DOT_EXACT   parallel order-reproducing sum: bit-identical to the serial loop (default)
 *   MFE_DOT_SERIAL  one device thread, the literal loop (validation of MFE_DOT_EXACT; slow)
 *   MFE_DOT_TREE    conventional warp/block tree (NOT reproducible vs the reference; comparison only) */
enum { MFE_DOT_EXACT = 0, MFE_DOT_SERIAL = 1, MFE_DOT_TREE = 2 };

int         mfe_abi_version(void);
const char *mfe_last_error(void);
int         mfe_device_count(void);

/* ref/utils/BoxPartition.cpp:79-122 as called at ref/src/main.cpp:112: box_partition(0,np,2,global,boxes).
 * boxes6n[6*p+{0..5}] = {x0,x1,y0,y1,z0,z1} element ranges of rank p.  Host only. */
int mfe_box_partition(int nx, int ny, int nz, int nranks, int *boxes6n);

/* Context for `rank` of `nranks` on CUDA device `device`.  global_box / all_boxes as produced by
 * mfe_box_partition (ref/src/main.cpp:100-114).  Replaces simple_mesh_description's ctor
 * (ref/src/simple_mesh_description.hpp:41-226): id->row map and BC sets become analytic functions of the
 * box table. */
mfe_ctx *mfe_create(int rank, int nranks, int device, const int global_box[6], const int *all_boxes6n);
void     mfe_destroy(mfe_ctx *ctx);

/* multi-GPU plumbing (replaces MPI_Init/MPI_COMM_WORLD): rank 0 calls mfe_comm_unique_id, the id is
 * distributed by the launcher (torch.distributed / file / MPI), every rank calls mfe_comm_init. */
int mfe_comm_unique_id(char id128[128]);
int mfe_comm_init(mfe_ctx *ctx, const char id128[128]);

/* ref/src/generate_matrix_structure.hpp:47-149 (+ MatrixInitOp.hpp:127-156, SparseMatrix_functions.hpp:56-74) */
int mfe_generate_matrix_structure(mfe_ctx *ctx, int64_t *nrows, int64_t *first_row, int64_t *nnz);
/* ref/src/assemble_FE_data.hpp:38-73 -> perform_element_loop.hpp:40-104 -> sum_into_global_linear_system
 * (SparseMatrix_functions.hpp:245-255).  b is zero-initialised by generate_matrix_structure (Vector.hpp:44-52). */
int mfe_assemble_FE_data(mfe_ctx *ctx);
/* ref/src/SparseMatrix_functions.hpp:431-480 impose_dirichlet(prescribed_value, A, b, ..., bc_rows):
 * which_set 0 = mesh.bc_rows_0 (faces x=0,y=0,z=0,y=1,z=1), 1 = mesh.bc_rows_1 (face x=1)
 * (ref/src/driver.hpp:215-218).  Must precede mfe_make_local_matrix (columns are still global rows). */
int mfe_impose_dirichlet(mfe_ctx *ctx, double prescribed_value, int which_set);
/* ref/src/make_local_matrix.hpp:42-442 */
int mfe_make_local_matrix(mfe_ctx *ctx, int64_t *num_cols);

/* ref/src/SparseMatrix_functions.hpp:319-408 rearrange_matrix_local_external, which the driver calls before cg_solve when
 * mv_overlap_comm_comp=1 (driver.hpp:278-282): every row is re-sorted so that its owned columns come first and its external
 * columns after them; the SpMV then sums each row exactly as the reference's matvec_overlap does (624-675: local entries
 * onto 0*y, external entries onto 1*y), which changes the rounding of the rows that have externals and hence the history. */
int mfe_rearrange_matrix_local_external(mfe_ctx *ctx);

/* ref/src/cg_solve.hpp:69-229 with matvec_std (SparseMatrix_functions.hpp:546-578).
 * x_host: in = initial guess, out = solution (nrows doubles; NULL: use/keep the device-resident x).
 * normr_hist (NULL or max_iter+1 doubles): hist[0] = initial residual norm, hist[k] = the norm the reference
 * would print for iteration k.  Returns 0, 1 on error, 2 on numerical breakdown (cg_solve.hpp:186-198:
 * the solve stops there, timers and outputs are still filled). */
int mfe_cg_solve(mfe_ctx *ctx, double *x_host, int max_iter, double tolerance, int *num_iters, double *normr,
                 double cg_times[MFE_NUM_TIMERS], double *normr_hist, double *rtrans_hist, double *pap_hist,
                 double *alpha_hist);
/* ref/src/time_kernels.hpp:42-134 (the MINIFE_KERNELS != 0 branch of driver.hpp:257-270): max_iter x waxpby(1,x,0,x,p),
 * max_iter x matvec(A,p,x), xdotp = sum over max_iter x dot(x,p), each loop timed (kern_times[WAXPY,MATVEC,DOT]).
 * x_host: in = x (nrows doubles), out = A*p as the reference leaves it; NULL = device-resident x. */
int mfe_time_kernels(mfe_ctx *ctx, double *x_host, int max_iter, double *xdotp, double kern_times[MFE_NUM_TIMERS]);
/* ref/fem/verify_solution.hpp:55-174 (+ analytic_soln.hpp:69-118) on the device-resident x left by mfe_cg_solve
 * (or mfe_import_vector "x"): compares x with the 301x301-term analytic series at the nodes within 0.05 of the
 * cube centre (verify_whole_domain = 0, driver.hpp:296) or at every owned node.  *return_code = the reference's
 * return value (1 when the GLOBAL max error exceeds tolerance); local_info[6] = this rank's {max err, computed,
 * analytic, x, y, z} (err_info, verify_solution.hpp:47-53); global_max_err = max over ranks. */
int mfe_verify_solution(mfe_ctx *ctx, double tolerance, int verify_whole_domain, int *return_code, double *global_max_err,
                        double local_info[6]);
int mfe_set_dot_mode(mfe_ctx *ctx, int mode);
/* knobs: "matvec_config" = index into the SpMV (warps, stages, gather batches) table; "comm_mode" = 1 NVLink peer
 * stores for halo + dot all-reduce (default when every rank can map every other rank's mailbox through CUDA IPC),
 * 0 NCCL send/recv + allgather; "overlap" = 1 run the SpMV tiles that need no halo while the halo is in flight
 * (peer mode); "peer_timeout_ms" = bound of every peer spin (default 30000); "dot_maps_warps" = 0 | 16 | 24; "fused" = 1 MINIFE_FUSED reporting (cg_solve.hpp:172-176,
 * 203-206): SpMV + dot(Ap,p) are timed as cg_times[MATVECDOT] (the values are bit-identical either way) */
int mfe_set_option(mfe_ctx *ctx, const char *name, int64_t value);
/* CUDA events inside mfe_cg_solve: 1 = around every op class, filling cg_times[WAXPY,DOT,MATVEC] (default);
 * 2 = around the SpMV only (two events per iteration: cg_times[MATVEC]); 0 = none (only TOTAL) */
int mfe_set_op_timing(mfe_ctx *ctx, int enabled);

/* kernel-level entry points on HOST buffers (H2D + kernel + D2H inside the call) -- used by parity tests and
 * by the reference-side shims of matvec_std / waxpby / dot.
 * mfe_matvec: ref matvec_std::operator() incl. exchange_externals; x has num_cols entries (externals are
 * overwritten by the halo exchange when nranks > 1), y has nrows. */
int mfe_matvec(mfe_ctx *ctx, double *x_host, double *y_host);
int mfe_waxpby(mfe_ctx *ctx, int64_t n, double alpha, const double *x_host, double beta, const double *y_host, double *w_host); /* Vector_functions.hpp:127-150 */
int mfe_dot(mfe_ctx *ctx, int64_t n, const double *x_host, const double *y_host, double *result); /* Vector_functions.hpp:194-226 (local part + rank-ordered sum) */
/* ref/src/exchange_externals.hpp:44-161 on a host vector of num_cols entries */
int mfe_exchange_externals(mfe_ctx *ctx, double *x_host);

/* sizes */
int64_t mfe_nrows(const mfe_ctx *ctx);
int64_t mfe_first_row(const mfe_ctx *ctx);
int64_t mfe_nnz(const mfe_ctx *ctx);
int64_t mfe_num_cols(const mfe_ctx *ctx);
/* name in {external_index, external_local_index, elements_to_send, neighbors, recv_length, send_length}
 * (ref/src/CSRMatrix.hpp:73-78); returns count (or -1), fills out[] when non-NULL */
int64_t mfe_get_map(const mfe_ctx *ctx, const char *name, int *out);

/* copy device state into caller-owned host buffers (any pointer may be NULL = skip).
 * rows[nrows], row_offsets[nrows+1] (int32 as CSRMatrix.hpp:64-66; fails if nnz > INT32_MAX, use
 * row_offsets64), packed_cols[nnz], packed_coefs[nnz] */
int mfe_export_csr(mfe_ctx *ctx, int *rows, int *row_offsets, int64_t *row_offsets64, int *packed_cols, double *packed_coefs);
int mfe_export_vector(mfe_ctx *ctx, const char *name /* "b" | "x" */, double *out);
int mfe_import_vector(mfe_ctx *ctx, const char *name /* "b" | "x" */, const double *in);
/* replace the device CSR by caller arrays already in local-index form (ref/basic utest pll_matvec style
 * hand matrices; also the shim path for a CSRMatrix assembled elsewhere).  int32 offsets. */
int mfe_import_csr(mfe_ctx *ctx, int64_t nrows, int64_t num_cols, const int *row_offsets, const int *packed_cols, const double *packed_coefs);

/* timed repetitions of single kernels on device-resident data, for bench.py / ncu: returns the average
 * milliseconds per launch over `reps` launches (CUDA events on the launching stream).
 * what: "matvec", "matvec_dot", "waxpby", "dot_rr", "dot_pap", "xr_update", "cg_iteration" */
int mfe_time_kernel(mfe_ctx *ctx, const char *what, int reps, double *ms_per_launch);
/* diagnostics of the order-reproducing dot: out[0] dots, [1] sum / [2] max of "dirty" leaves (replayed term by term),
 * [3] whole-vector replays, [4] gaps that took the slow path, [5] total / [6] max ns in the walk kernel, [7] index of
 * the slowest dot; from out[16]: a ring of 512 (slow gaps << 32 | dirty leaves, ns) pairs.  Returns the word count. */
int64_t mfe_dot_stats(mfe_ctx *ctx, uint64_t *out, int64_t cap, int reset);
/* number of kernel launches issued by this context since creation (bench.py's gpu_launches) */
int64_t mfe_launch_count(const mfe_ctx *ctx);
/* phase timings of the last calls, milliseconds: "structure","assembly","assembly_elem","assembly_gather",
 * "dirichlet0","dirichlet1","make_local","cg" */
double mfe_phase_ms(const mfe_ctx *ctx, const char *phase);

#ifdef __cplusplus
}
#endif
#endif
