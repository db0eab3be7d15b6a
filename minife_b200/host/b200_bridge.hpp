// b200_bridge.hpp -- glue between the reference's host types (miniFE::CSRMatrix / Vector / simple_mesh_description,
// used UNCHANGED from the reference tree) and the C ABI of libminife_b200.so.
//
// The reference's operators are free function templates called by ref/src/driver.hpp:124-404.  The headers in this
// directory carry the same names and signatures and are put FIRST on the include path, so the unmodified main.cpp
// and driver.hpp compile against them; each shim marshals to one mfe_* call.  Device state is kept in a side table
// keyed by the address of the host object (precedent in the reference: ref/basic/optional/cuda/CudaMemoryModel.hpp),
// because the host structs must stay exactly as the reference declares them.
#ifndef MINIFE_B200_BRIDGE_HPP
#define MINIFE_B200_BRIDGE_HPP

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef HAVE_MPI
#include <mpi.h>
#endif

#include <minife_b200.h>
#include <Box.hpp>

namespace miniFE {
namespace b200 {

struct Entry {
  mfe_ctx *ctx;
  const void *bc_rows_0, *bc_rows_1;   // addresses of the mesh's own BC sets (driver.hpp:215-218 passes them back)
  const void *b_vec, *x_vec;           // host Vectors bound to the device b / x
  Entry() : ctx(0), bc_rows_0(0), bc_rows_1(0), b_vec(0), x_vec(0) {}
};

inline std::map<const void *, Entry> &table()
{
  static std::map<const void *, Entry> t;
  return t;
}

inline void die(const char *what)
{
  std::cerr << "minife_b200: " << what << " failed: " << mfe_last_error() << std::endl;
#ifdef HAVE_MPI
  MPI_Abort(MPI_COMM_WORLD, -1);
#endif
  std::exit(1);
}

inline Entry &entry_of(const void *A, const char *who)
{
  std::map<const void *, Entry>::iterator it = table().find(A);
  if (it == table().end() || !it->second.ctx) {
    std::cerr << "minife_b200: " << who << ": matrix was not created by generate_matrix_structure" << std::endl;
    std::exit(1);
  }
  return it->second;
}

inline Entry *entry_of_vector(const void *v)
{
  for (std::map<const void *, Entry>::iterator it = table().begin(); it != table().end(); ++it)
    if (it->second.b_vec == v || it->second.x_vec == v) return &it->second;
  return 0;
}

// The MINIFE_DEBUG dump files (write_matrix / write_vector, ref/src/SparseMatrix_functions.hpp:98-140 and
// ref/src/Vector_functions.hpp:49-82): one text file per rank named "<base>.<numprocs>.<rank>", written by the ranks
// one after the other; only rank 0 writes a header line.  `body(os, is_rank0)` emits this rank's lines.
struct RankFile {
  int numprocs, myproc;
  std::ofstream os;
  explicit RankFile(const std::string &base) : numprocs(1), myproc(0)
  {
#ifdef HAVE_MPI
    MPI_Comm_size(MPI_COMM_WORLD, &numprocs);
    MPI_Comm_rank(MPI_COMM_WORLD, &myproc);
#endif
    std::ostringstream name;
    name << base << '.' << numprocs << '.' << myproc;
    os.open(name.str().c_str());
  }
  template<typename Body>
  void in_rank_order(Body body)
  {
    for (int turn = 0; turn < numprocs; ++turn) {
      if (turn == myproc) body(os, myproc == 0);
#ifdef HAVE_MPI
      MPI_Barrier(MPI_COMM_WORLD);
#endif
    }
  }
};

// one context per rank: rank/size from MPI (1/0 without), boxes gathered from all ranks (they may have been
// changed by add_imbalance, driver.hpp:137-143), GPU = rank modulo the visible device count
inline mfe_ctx *create_context(const Box &global_box, const Box &local_box)
{
  int numprocs = 1, myproc = 0;
#ifdef HAVE_MPI
  MPI_Comm_size(MPI_COMM_WORLD, &numprocs);
  MPI_Comm_rank(MPI_COMM_WORLD, &myproc);
#endif
  std::vector<int> all(6 * numprocs);
  for (int i = 0; i < 6; ++i) all[6 * myproc + i] = local_box.ranges[i];
#ifdef HAVE_MPI
  std::vector<int> mine(local_box.ranges, local_box.ranges + 6);
  MPI_Allgather(&mine[0], 6, MPI_INT, &all[0], 6, MPI_INT, MPI_COMM_WORLD);
#endif
  int ndev = mfe_device_count();
  if (ndev < 1) { std::cerr << "minife_b200: no CUDA device; this build has no CPU path" << std::endl; std::exit(1); }
  mfe_ctx *ctx = mfe_create(myproc, numprocs, myproc % ndev, global_box.ranges, &all[0]);
  if (!ctx) die("mfe_create");
  if (numprocs > 1) {
    char id[128];
    if (myproc == 0 && mfe_comm_unique_id(id) != 0) die("mfe_comm_unique_id");
#ifdef HAVE_MPI
    MPI_Bcast(id, 128, MPI_CHAR, 0, MPI_COMM_WORLD);
#endif
    if (mfe_comm_init(ctx, id) != 0) die("mfe_comm_init");
  }
  return ctx;
}

}  // namespace b200
}  // namespace miniFE

#endif
