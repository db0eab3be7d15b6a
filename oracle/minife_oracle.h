/* oracle/minife_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the miniFE `ref/` hot path (hex8 assembly + unpreconditioned CG),
 * written from the reference's algorithm, each function citing the reference file:line it follows
 * (paths relative to /root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product (minife_b200/) never
 * links, loads or calls it.
 *
 * PARITY PINNED: tests/test_oracle_vs_ref.py checks every function here bit-for-bit against the
 * unmodified reference headers (oracle/_ref/libminife_ref.so, built from /root/reference by
 * oracle/Makefile) and against the committed fixtures under tests/golden/ generated from it.
 *
 * All ranks of a run live in ONE process ("world"); MPI messages of the reference become plain
 * array reads between the rank structs, allreduces are summed in rank order ((r0+r1)+r2)+...,
 * which is the association oracle/mpi_shim/mpi.h gives the unmodified reference.
 *
 * Build: gcc -O2 -ffp-contract=off (no -march, no -ffast-math): one rounding per operation.
 */
#ifndef MINIFE_ORACLE_H
#define MINIFE_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mfo_world mfo_world;

/* ref/utils/BoxPartition.cpp:79-122  box_partition(0, np, 2, global, out) */
void mfo_box_partition(int nx, int ny, int nz, int np, int *out6np);

/* ref/fem/Hex8_ElemData.hpp:37-61 (192 gradient values at the 8 Gauss points) */
void mfo_gradient_values(double *out192);
/* ref/src/Hex8_box_utils.hpp:101-159 + ref/fem/Hex8_ElemData.hpp:63-71 for one GLOBAL element id of an
 * nx*ny*nz mesh: node ids (exodus order, node ids not rows), 24 coords, 36 K, 8 f */
void mfo_element(int nx, int ny, int nz, int elemID, int *node_ids8, double *coords24, double *K36, double *f8);

mfo_world *mfo_world_create(int nx, int ny, int nz, int nranks);
/* same with caller-supplied element boxes (6 ints per rank), e.g. after add_imbalance (ref/utils/imbalance.hpp:167-293) */
mfo_world *mfo_world_create_boxes(int nx, int ny, int nz, int nranks, const int *boxes6n);
void mfo_world_destroy(mfo_world *w);

/* phases (all ranks), in the order of ref/src/driver.hpp:159-289 */
int  mfo_generate_matrix_structure(mfo_world *w);     /* generate_matrix_structure.hpp:47-149 */
void mfo_assemble_FE_data(mfo_world *w);              /* assemble_FE_data.hpp:38-73, perform_element_loop.hpp:40-104 */
void mfo_impose_dirichlet(mfo_world *w, int which);   /* SparseMatrix_functions.hpp:431-480; which 0: (0.0,bc_rows_0) 1: (1.0,bc_rows_1) */
void mfo_make_local_matrix(mfo_world *w);             /* make_local_matrix.hpp:42-442 */
/* cg_solve.hpp:69-229 with matvec_std; histories may be NULL; hist arrays hold max_iter+1 entries */
int  mfo_cg_solve(mfo_world *w, int max_iter, double tol, int *num_iters, double *normr,
                  double *normr_hist, double *rtrans_hist, double *pap_hist, double *alpha_hist);

/* ref/fem/verify_solution.hpp:55-174 + analytic_soln.hpp:56-118 on the x of every rank (as left by mfo_cg_solve or
 * mfo_set_x).  Returns the reference's return code; info6 = {max err, computed, analytic, x, y, z}. */
int  mfo_verify_solution(mfo_world *w, double tolerance, int verify_whole_domain, double *global_max_err, double *info6);

/* per-rank queries */
long long mfo_nrows(mfo_world *w, int rank);
long long mfo_nnz(mfo_world *w, int rank);
long long mfo_num_cols(mfo_world *w, int rank);
long long mfo_first_row(mfo_world *w, int rank);
void mfo_get_box(mfo_world *w, int rank, int *box6);
void mfo_get_csr(mfo_world *w, int rank, int *rows, int *row_offsets, int *cols, double *coefs);
void mfo_get_b(mfo_world *w, int rank, double *out);
void mfo_get_x(mfo_world *w, int rank, double *out);
void mfo_set_x(mfo_world *w, int rank, const double *in);
long long mfo_bc_rows(mfo_world *w, int rank, int which, int *out);   /* ascending; returns count */
int  mfo_map_id_to_row(mfo_world *w, int rank, int id);
/* index maps of make_local_matrix; name in {external_index, external_local_index, elements_to_send,
 * neighbors, recv_length, send_length}; returns count, fills out when non-NULL */
long long mfo_get_map(mfo_world *w, int rank, const char *name, int *out);

/* kernel-level entry points on caller arrays (single rank semantics) */
/* SparseMatrix_functions.hpp:546-578 without the halo exchange: x has num_cols entries */
void mfo_matvec(mfo_world *w, int rank, const double *x, double *y);
/* exchange_externals.hpp:44-161 on per-rank arrays x[rank] (num_cols entries each) */
void mfo_exchange_externals(mfo_world *w, double **x_per_rank);
void mfo_waxpby(long long n, double alpha, const double *x, double beta, const double *y, double *w); /* Vector_functions.hpp:127-150 */
double mfo_dot(long long n, const double *x, const double *y);                                       /* Vector_functions.hpp:194-226 */

/* ref/src/SparseMatrix_functions.hpp:319-408 (+ 624-675): after this, mfo_matvec / mfo_cg_solve sum every row in the order
 * of the reference's matvec_overlap (local entries first, then externals; driver option mv_overlap_comm_comp=1) */
void mfo_rearrange_matrix_local_external(mfo_world *w);

#ifdef __cplusplus
}
#endif
#endif
