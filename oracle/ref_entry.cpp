// oracle/ref_entry.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Binds the UNMODIFIED reference headers (included straight from
// /root/reference/ref/{src,utils,fem}; nothing is copied into this repo) behind
// a small extern "C" surface so that tests can ask "what does the reference
// itself compute?" in full binary precision.  Built by oracle/Makefile into
// oracle/_ref/libminife_ref.so (git-ignored; travels to the GPU box).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
// the result.  The product (minife_b200/) never links or loads it.
//
// With -DHAVE_MPI -Ioracle/mpi_shim the same TU builds ref_dump_mpi.x, a
// per-rank process that dumps the reference's multi-rank structure / index maps
// / values / residual history as raw binary files (see REF_DUMP_MAIN below).
#include <iostream>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include <limits>
#include <cstring>
#include <cstdio>
#include <cstdlib>

#ifdef HAVE_MPI
#include <mpi.h>
#endif

#include <Box.hpp>
#include <BoxPartition.hpp>
#include <box_utils.hpp>
#include <Parameters.hpp>
#include <utils.hpp>
#include <mytimer.hpp>
#include <Vector.hpp>
#include <CSRMatrix.hpp>
#include <simple_mesh_description.hpp>
#include <SparseMatrix_functions.hpp>
#include <generate_matrix_structure.hpp>
#include <assemble_FE_data.hpp>
#include <make_local_matrix.hpp>
#include <Vector_functions.hpp>
#include <cg_solve.hpp>
#include <verify_solution.hpp>
#ifdef REF_DUMP_MAIN
#include <YAML_Doc.hpp>
#include <imbalance.hpp>
#endif

typedef miniFE::CSRMatrix<double, int, int> RefMatrix;
typedef miniFE::Vector<double, int, int> RefVector;
typedef miniFE::simple_mesh_description<int> RefMesh;

namespace {

// Silences the reference's chatter on std::cout for the lifetime of the object
// and keeps what was printed (the CG residual lines are parsed from it).
struct CoutCapture {
  std::ostringstream sink;
  std::streambuf* saved;
  std::streamsize saved_prec;
  explicit CoutCapture(int precision = 17) {
    saved = std::cout.rdbuf(sink.rdbuf());
    saved_prec = std::cout.precision(precision);
  }
  ~CoutCapture() {
    std::cout.rdbuf(saved);
    std::cout.precision(saved_prec);
  }
};

struct RefProblem {
  Box global_box;
  Box my_box;
  int numprocs, myproc;
  RefMesh* mesh;
  RefMatrix A;
  RefVector* b;
  RefVector* x;
  miniFE::Parameters params;
  RefProblem() : mesh(0), b(0), x(0) {}
  ~RefProblem() { delete mesh; delete b; delete x; }
};

RefProblem* make_problem(int nx, int ny, int nz, int numprocs, int myproc)
{
  RefProblem* P = new RefProblem;
  P->numprocs = numprocs;
  P->myproc = myproc;
  Box g = { 0, nx, 0, ny, 0, nz };
  P->global_box = g;
  std::vector<Box> boxes(numprocs);
  box_partition(0, numprocs, 2, g, &boxes[0]);   // ref/src/main.cpp:112
  P->my_box = boxes[myproc];
  P->params.nx = nx; P->params.ny = ny; P->params.nz = nz;
  CoutCapture quiet;
  P->mesh = new RefMesh(P->global_box, P->my_box);
  return P;
}

void make_vectors(RefProblem* P)
{
  int local_nrows = P->A.rows.size();
  int first = local_nrows > 0 ? P->A.rows[0] : -1;   // ref/src/driver.hpp:185-189
  delete P->b; delete P->x;
  P->b = new RefVector(first, local_nrows);
  P->x = new RefVector(first, local_nrows);
}

} // namespace

extern "C" {

void* ref_create(int nx, int ny, int nz) { return make_problem(nx, ny, nz, 1, 0); }
void ref_destroy(void* h) { delete static_cast<RefProblem*>(h); }

int ref_generate_matrix_structure(void* h)
{
  RefProblem* P = static_cast<RefProblem*>(h);
  CoutCapture quiet;
  int rc = miniFE::generate_matrix_structure(*P->mesh, P->A);
  make_vectors(P);
  return rc;
}

void ref_assemble_FE_data(void* h)
{
  RefProblem* P = static_cast<RefProblem*>(h);
  CoutCapture quiet;
  miniFE::assemble_FE_data(*P->mesh, P->A, *P->b, P->params);
}

// which: 0 -> impose_dirichlet(0.0, bc_rows_0), 1 -> impose_dirichlet(1.0, bc_rows_1)
void ref_impose_dirichlet(void* h, int which)
{
  RefProblem* P = static_cast<RefProblem*>(h);
  CoutCapture quiet;
  int gx = P->global_box[0][1] + 1, gy = P->global_box[1][1] + 1, gz = P->global_box[2][1] + 1;
  if (which == 0) miniFE::impose_dirichlet(0.0, P->A, *P->b, gx, gy, gz, P->mesh->bc_rows_0);
  else            miniFE::impose_dirichlet(1.0, P->A, *P->b, gx, gy, gz, P->mesh->bc_rows_1);
}

void ref_make_local_matrix(void* h)
{
  RefProblem* P = static_cast<RefProblem*>(h);
  CoutCapture quiet;
  miniFE::make_local_matrix(P->A);
}

long long ref_nrows(void* h) { return static_cast<RefProblem*>(h)->A.rows.size(); }
long long ref_nnz(void* h)   { return static_cast<RefProblem*>(h)->A.num_nonzeros(); }
long long ref_num_cols(void* h) { return static_cast<RefProblem*>(h)->A.num_cols; }
long long ref_first_row(void* h)
{
  RefProblem* P = static_cast<RefProblem*>(h);
  return P->A.rows.empty() ? -1 : P->A.rows[0];
}

void ref_get_csr(void* h, int* rows, int* row_offsets, int* cols, double* coefs)
{
  RefProblem* P = static_cast<RefProblem*>(h);
  size_t n = P->A.rows.size(), nnz = P->A.num_nonzeros();
  if (rows) std::memcpy(rows, &P->A.rows[0], n * sizeof(int));
  if (row_offsets) std::memcpy(row_offsets, &P->A.row_offsets[0], (n + 1) * sizeof(int));
  if (cols) std::memcpy(cols, &P->A.packed_cols[0], nnz * sizeof(int));
  if (coefs) std::memcpy(coefs, &P->A.packed_coefs[0], nnz * sizeof(double));
}

void ref_get_b(void* h, double* out)
{
  RefProblem* P = static_cast<RefProblem*>(h);
  std::memcpy(out, &P->b->coefs[0], P->b->coefs.size() * sizeof(double));
}
void ref_get_x(void* h, double* out)
{
  RefProblem* P = static_cast<RefProblem*>(h);
  std::memcpy(out, &P->x->coefs[0], P->x->coefs.size() * sizeof(double));
}
void ref_set_x(void* h, const double* in)
{
  RefProblem* P = static_cast<RefProblem*>(h);
  std::memcpy(&P->x->coefs[0], in, P->x->coefs.size() * sizeof(double));
}

// Sizes then contents of bc_rows_0 / bc_rows_1 (ascending, std::set order).
long long ref_bc_rows(void* h, int which, int* out)
{
  RefProblem* P = static_cast<RefProblem*>(h);
  const std::set<int>& s = which == 0 ? P->mesh->bc_rows_0 : P->mesh->bc_rows_1;
  if (out) { size_t i = 0; for (std::set<int>::const_iterator it = s.begin(); it != s.end(); ++it) out[i++] = *it; }
  return (long long)s.size();
}

int ref_map_id_to_row(void* h, int id) { return static_cast<RefProblem*>(h)->mesh->map_id_to_row(id); }

// One element through get_elem_nodes_and_coords + compute_element_matrix_and_vector
// (ref/src/Hex8_box_utils.hpp:101-159, ref/fem/Hex8_ElemData.hpp:63-71).
void ref_element(void* h, int elemID, int* node_rows, double* coords, double* K36, double* f8)
{
  RefProblem* P = static_cast<RefProblem*>(h);
  miniFE::ElemData<int, double> ed;
  miniFE::compute_gradient_values(ed.grad_vals);
  miniFE::get_elem_nodes_and_coords(*P->mesh, elemID, ed);
  miniFE::compute_element_matrix_and_vector(ed);
  if (node_rows) std::memcpy(node_rows, ed.elem_node_ids, 8 * sizeof(int));
  if (coords) std::memcpy(coords, ed.elem_node_coords, 24 * sizeof(double));
  if (K36) std::memcpy(K36, ed.elem_diffusion_matrix, 36 * sizeof(double));
  if (f8) std::memcpy(f8, ed.elem_source_vector, 8 * sizeof(double));
}

void ref_gradient_values(double* out192)
{
  miniFE::compute_gradient_values(out192);
}

// Kernel-level entry points on caller arrays (copied through reference Vectors).
void ref_matvec(void* h, const double* xin, double* yout)
{
  RefProblem* P = static_cast<RefProblem*>(h);
  int nrows = P->A.rows.size(), ncols = P->A.num_cols;
  RefVector xv(0, ncols), yv(0, nrows);
  std::memcpy(&xv.coefs[0], xin, ncols * sizeof(double));
  std::memcpy(&yv.coefs[0], yout, nrows * sizeof(double));   // matvec reads 0*y (SparseMatrix_functions.hpp:565)
  miniFE::matvec_std<RefMatrix, RefVector> mv;
  mv(P->A, xv, yv);
  std::memcpy(yout, &yv.coefs[0], nrows * sizeof(double));
}

void ref_waxpby(int n, double alpha, const double* x, double beta, const double* y, double* w)
{
  RefVector xv(0, n), yv(0, n), wv(0, n);
  std::memcpy(&xv.coefs[0], x, n * sizeof(double));
  std::memcpy(&yv.coefs[0], y, n * sizeof(double));
  miniFE::waxpby(alpha, xv, beta, yv, wv);
  std::memcpy(w, &wv.coefs[0], n * sizeof(double));
}

double ref_dot(int n, const double* x, const double* y)
{
  RefVector xv(0, n), yv(0, n);
  std::memcpy(&xv.coefs[0], x, n * sizeof(double));
  std::memcpy(&yv.coefs[0], y, n * sizeof(double));
  return miniFE::dot(xv, yv);
}

// The unmodified cg_solve (ref/src/cg_solve.hpp:69-229).  printed[k] receives the
// residual the reference printed for iteration k (k=0: "Initial Residual"), parsed
// at 17 significant digits; entries it did not print are left untouched.
int ref_cg_solve(void* h, int max_iter, double tol, int* num_iters, double* normr,
                 double* cg_times, double* printed, int printed_len)
{
  RefProblem* P = static_cast<RefProblem*>(h);
  miniFE::timer_type times[miniFE::NUM_TIMERS] = {0, 0, 0, 0, 0};
  int iters = 0;
  double rnorm = 0;
  std::string text;
  {
    CoutCapture cap(17);
    miniFE::cg_solve(P->A, *P->b, *P->x, miniFE::matvec_std<RefMatrix, RefVector>(),
                     max_iter, tol, iters, rnorm, times);
    text = cap.sink.str();
  }
  *num_iters = iters;
  *normr = rnorm;
  if (cg_times) for (int i = 0; i < miniFE::NUM_TIMERS; ++i) cg_times[i] = times[i];
  if (printed) {
    std::istringstream in(text);
    std::string line;
    while (std::getline(in, line)) {
      int k; double r;
      if (std::sscanf(line.c_str(), "Initial Residual = %lf", &r) == 1) { if (printed_len > 0) printed[0] = r; }
      else if (std::sscanf(line.c_str(), "Iteration = %d   Residual = %lf", &k, &r) == 2) { if (k < printed_len) printed[k] = r; }
    }
  }
  return 0;
}

// The same iteration written out with the reference's OWN matvec_std / dot / waxpby so
// that every per-iteration scalar is observable.  tests/ check that it ends bit-identical
// to ref_cg_solve.  Sequence follows ref/src/cg_solve.hpp:124-215.
// normr_hist[k] (k=0..num_iters): residual norm as the reference would report for iteration k.
static int g_trace_overlap = 0;   // 1: the trace uses the reference's matvec_overlap (after rearrange_matrix_local_external)
struct TraceMatvec {
  void operator()(RefMatrix& A, RefVector& x, RefVector& y) {
    if (g_trace_overlap) { miniFE::matvec_overlap<RefMatrix, RefVector> mv; mv(A, x, y); }
    else { miniFE::matvec_std<RefMatrix, RefVector> mv; mv(A, x, y); }
  }
};
void ref_rearrange_matrix_local_external(void* h)
{
  RefProblem* P = static_cast<RefProblem*>(h);
  miniFE::rearrange_matrix_local_external(P->A);
  g_trace_overlap = 1;
}

int ref_cg_trace(void* h, int max_iter, double tol, int* num_iters, double* normr_out,
                 double* normr_hist, double* rtrans_hist, double* pap_hist, double* alpha_hist)
{
  RefProblem* P = static_cast<RefProblem*>(h);
  RefMatrix& A = P->A;
  RefVector& b = *P->b;
  RefVector& x = *P->x;
  size_t nrows = A.rows.size();
  RefVector r(b.startIndex, nrows), p(0, A.num_cols), Ap(b.startIndex, nrows);
  TraceMatvec mv;
  double one = 1.0, zero = 0.0;
  miniFE::waxpby(one, x, zero, x, p);
  mv(A, p, Ap);
  miniFE::waxpby(one, b, -one, Ap, r);
  double rtrans = miniFE::dot(r, r), oldrtrans = 0;
  double normr = std::sqrt(rtrans);
  if (normr_hist) normr_hist[0] = normr;
  if (rtrans_hist) rtrans_hist[0] = rtrans;
  double brkdown_tol = std::numeric_limits<double>::epsilon();
  int iters = 0, rc = 0;
  for (int k = 1; k <= max_iter && normr > tol; ++k) {
    if (k == 1) miniFE::waxpby(one, r, zero, r, p);
    else {
      oldrtrans = rtrans;
      rtrans = miniFE::dot(r, r);
      double beta = rtrans / oldrtrans;
      miniFE::waxpby(one, r, beta, p, p);
    }
    normr = std::sqrt(rtrans);
    if (normr_hist) normr_hist[k] = normr;
    if (rtrans_hist) rtrans_hist[k] = rtrans;
    mv(A, p, Ap);
    double p_ap_dot = miniFE::dot(Ap, p);
    if (pap_hist) pap_hist[k] = p_ap_dot;
    if (p_ap_dot < brkdown_tol) {
      if (p_ap_dot < 0 || miniFE::breakdown(p_ap_dot, Ap, p)) { rc = 1; break; }
      else brkdown_tol = 0.1 * p_ap_dot;
    }
    double alpha = rtrans / p_ap_dot;
    if (alpha_hist) alpha_hist[k] = alpha;
    miniFE::waxpby(one, x, alpha, p, x);
    miniFE::waxpby(one, r, -alpha, Ap, r);
    iters = k;
  }
  *num_iters = iters;
  *normr_out = normr;
  return rc;
}

int ref_verify_solution(void* h, double tolerance, int whole_domain)
{
  RefProblem* P = static_cast<RefProblem*>(h);
  CoutCapture quiet;
  return miniFE::verify_solution(*P->mesh, *P->x, tolerance, whole_domain != 0);
}

// box_partition(0, np, 2, global, out) -- ref/utils/BoxPartition.cpp:79-122
void ref_box_partition(int nx, int ny, int nz, int np, int* out6np)
{
  Box g = { 0, nx, 0, ny, 0, nz };
  std::vector<Box> boxes(np);
  box_partition(0, np, 2, g, &boxes[0]);
  for (int p = 0; p < np; ++p) for (int k = 0; k < 6; ++k) out6np[6 * p + k] = boxes[p].ranges[k];
}

} // extern "C"

#ifdef REF_DUMP_MAIN
// ---------------------------------------------------------------------------
// Multi-rank dump process: ref_dump_mpi.x nx ny nz outdir [cg_max_iter [load_imbalance_percent]]
// (load_imbalance > 0: the boxes are changed by add_imbalance exactly as ref/src/driver.hpp:137-139 does)
// Runs the reference pipeline on this rank (MPI via oracle/mpi_shim) and writes
//   <outdir>/<name>.<np>.<rank>.{i32,f64}
// ---------------------------------------------------------------------------
namespace {
template <typename T>
void dump(const std::string& dir, const char* name, int np, int rank, const char* ext, const T* data, size_t n)
{
  std::ostringstream fn;
  fn << dir << "/" << name << "." << np << "." << rank << "." << ext;
  std::ofstream f(fn.str().c_str(), std::ios::binary);
  f.write(reinterpret_cast<const char*>(data), n * sizeof(T));
}
template <typename T>
void dumpv(const std::string& dir, const char* name, int np, int rank, const char* ext, const std::vector<T>& v)
{
  dump(dir, name, np, rank, ext, v.empty() ? (const T*)0 : &v[0], v.size());
}
// CRC-32 (IEEE 802.3 / zlib polynomial, reflected) -- equals Python's zlib.crc32 over the array's bytes
unsigned crc32_bytes(const unsigned char* p, size_t n)
{
  static unsigned table[8][256];
  static bool init = false;
  if (!init) {
    for (unsigned i = 0; i < 256; ++i) {
      unsigned c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      table[0][i] = c;
    }
    for (unsigned i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) table[t][i] = (table[t - 1][i] >> 8) ^ table[0][table[t - 1][i] & 0xff];
    init = true;
  }
  unsigned c = 0xFFFFFFFFu;
  while (n >= 8) {                      // slicing-by-8
    unsigned lo, hi;
    std::memcpy(&lo, p, 4); std::memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = table[7][lo & 0xff] ^ table[6][(lo >> 8) & 0xff] ^ table[5][(lo >> 16) & 0xff] ^ table[4][lo >> 24] ^
        table[3][hi & 0xff] ^ table[2][(hi >> 8) & 0xff] ^ table[1][(hi >> 16) & 0xff] ^ table[0][hi >> 24];
    p += 8; n -= 8;
  }
  while (n--) c = table[0][(c ^ *p++) & 0xff] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}
template <typename T>
unsigned crc_of(const std::vector<T>& v)
{
  return crc32_bytes(v.empty() ? (const unsigned char*)0 : reinterpret_cast<const unsigned char*>(&v[0]), v.size() * sizeof(T));
}
}

int main(int argc, char** argv)
{
  int numprocs = 1, myproc = 0;
  miniFE::initialize_mpi(argc, argv, numprocs, myproc);
  if (argc < 5) { std::cerr << "usage: ref_dump_mpi.x nx ny nz outdir [max_iter]\n"; return 2; }
  int nx = std::atoi(argv[1]), ny = std::atoi(argv[2]), nz = std::atoi(argv[3]);
  std::string dir = argv[4];
  int max_iter = argc > 5 ? std::atoi(argv[5]) : 200;
  float load_imbalance = argc > 6 ? (float)std::atof(argv[6]) : 0.0f;

  // REF_DUMP_LEAN=1: large meshes -- no array dumps; CRC-32 (zlib polynomial) of every array instead, the index maps,
  // and the four per-iteration histories.  Same pipeline, same order of calls.
  const bool lean = std::getenv("REF_DUMP_LEAN") != 0 && std::atoi(std::getenv("REF_DUMP_LEAN")) != 0;
  std::vector<unsigned> crcs;

  RefProblem* P = make_problem(nx, ny, nz, numprocs, myproc);
  if (load_imbalance > 0) {      // ref/src/driver.hpp:137-139, then the mesh is (re)built from the changed box
    YAML_Doc ydoc("miniFE", "dump");
    { CoutCapture quiet; miniFE::add_imbalance<int>(P->global_box, P->my_box, load_imbalance, ydoc); }
    delete P->mesh;
    CoutCapture quiet;
    P->mesh = new RefMesh(P->global_box, P->my_box);
  }
  { CoutCapture quiet; miniFE::generate_matrix_structure(*P->mesh, P->A); }
  make_vectors(P);
  dump(dir, "box", numprocs, myproc, "i32", P->my_box.ranges, 6);
  if (lean) {
    crcs.push_back(crc_of(P->A.row_offsets)); crcs.push_back(crc_of(P->A.packed_cols));
  } else {
    dumpv(dir, "rows", numprocs, myproc, "i32", P->A.rows);
    dumpv(dir, "row_offsets", numprocs, myproc, "i32", P->A.row_offsets);
    dumpv(dir, "cols_global", numprocs, myproc, "i32", P->A.packed_cols);
  }
  ref_assemble_FE_data(P);
  if (lean) { crcs.push_back(crc_of(P->A.packed_coefs)); crcs.push_back(crc_of(P->b->coefs)); }
  else {
    dumpv(dir, "coefs_prebc", numprocs, myproc, "f64", P->A.packed_coefs);
    dumpv(dir, "b_prebc", numprocs, myproc, "f64", P->b->coefs);
  }
  ref_impose_dirichlet(P, 0);
  ref_impose_dirichlet(P, 1);
  if (lean) { crcs.push_back(crc_of(P->A.packed_coefs)); crcs.push_back(crc_of(P->b->coefs)); }
  else {
    dumpv(dir, "coefs", numprocs, myproc, "f64", P->A.packed_coefs);
    dumpv(dir, "b", numprocs, myproc, "f64", P->b->coefs);
  }
  ref_make_local_matrix(P);
  if (lean) crcs.push_back(crc_of(P->A.packed_cols));
  else dumpv(dir, "cols_local", numprocs, myproc, "i32", P->A.packed_cols);
  int num_cols = P->A.num_cols;
  dump(dir, "num_cols", numprocs, myproc, "i32", &num_cols, 1);
  long long sizes[2] = { (long long)P->A.rows.size(), (long long)P->A.packed_cols.size() };
  dump(dir, "sizes", numprocs, myproc, "i64", sizes, 2);
#ifdef HAVE_MPI
  dumpv(dir, "external_index", numprocs, myproc, "i32", P->A.external_index);
  dumpv(dir, "external_local_index", numprocs, myproc, "i32", P->A.external_local_index);
  dumpv(dir, "elements_to_send", numprocs, myproc, "i32", P->A.elements_to_send);
  dumpv(dir, "neighbors", numprocs, myproc, "i32", P->A.neighbors);
  dumpv(dir, "recv_length", numprocs, myproc, "i32", P->A.recv_length);
  dumpv(dir, "send_length", numprocs, myproc, "i32", P->A.send_length);
#endif
  // REF_DUMP_OVERLAP=1: the driver's mv_overlap_comm_comp=1 branch (driver.hpp:278-282)
  if (std::getenv("REF_DUMP_OVERLAP") != 0 && std::atoi(std::getenv("REF_DUMP_OVERLAP")) != 0) ref_rearrange_matrix_local_external(P);
  if (max_iter > 0) {
    std::vector<double> nh(max_iter + 1, -1.0), rh(max_iter + 1, -1.0), ph(max_iter + 1, -1.0), ah(max_iter + 1, -1.0);
    int iters = 0; double normr = 0;
    double tol = std::numeric_limits<double>::epsilon();
    ref_cg_trace(P, max_iter, tol, &iters, &normr, &nh[0], &rh[0], &ph[0], &ah[0]);
    nh.resize(iters + 1); rh.resize(iters + 1); ph.resize(iters + 1); ah.resize(iters + 1);
    dumpv(dir, "normr_hist", numprocs, myproc, "f64", nh);
    dumpv(dir, "pap_hist", numprocs, myproc, "f64", ph);
    dumpv(dir, "rtrans_hist", numprocs, myproc, "f64", rh);
    dumpv(dir, "alpha_hist", numprocs, myproc, "f64", ah);
    if (lean) crcs.push_back(crc_of(P->x->coefs));
    else dumpv(dir, "x", numprocs, myproc, "f64", P->x->coefs);
    if (myproc == 0) std::cout << "np=" << numprocs << " iters=" << iters << " normr=" << normr << std::endl;
  }
  // order: row_offsets, cols_global, coefs_prebc, b_prebc, coefs, b, cols_local, x
  if (lean) dumpv(dir, "crc", numprocs, myproc, "u32", crcs);
  delete P;
  miniFE::finalize_mpi();
  return 0;
}
#endif
