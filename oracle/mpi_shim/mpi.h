// oracle/mpi_shim/mpi.h -- TEST INFRASTRUCTURE.
//
// A minimal single-host stand-in for the 15 MPI entry points the reference's
// ref/src + ref/utils + ref/fem use, so that the
// UNMODIFIED reference can run with np > 1 in an image that has no MPI.  One
// process per rank; every pair of ranks shares one AF_UNIX socketpair whose
// descriptors are handed down by oracle/mpi_shim/launch.py through the
// MPISHIM_RANK / MPISHIM_SIZE / MPISHIM_FDS environment variables.
//
// Reductions are evaluated on rank 0 in rank order ((r0 + r1) + r2) + ... and
// broadcast; this fixes the association of the reference's dot-product
// allreduce, which the product's rank-ordered device reduction reproduces.
#ifndef MINIFE_B200_MPI_SHIM_H
#define MINIFE_B200_MPI_SHIM_H

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cerrno>
#include <deque>
#include <vector>
#include <string>
#include <unistd.h>
#include <fcntl.h>
#include <poll.h>
#include <sys/time.h>
#include <sys/socket.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
struct MPI_Status { int MPI_SOURCE; int MPI_TAG; int MPI_ERROR; };

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)

enum { MPI_CHAR = 1, MPI_INT, MPI_LONG, MPI_LONG_LONG, MPI_UNSIGNED, MPI_UNSIGNED_LONG,
       MPI_FLOAT, MPI_DOUBLE, MPI_COMPLEX, MPI_DOUBLE_COMPLEX, MPI_BYTE };
enum { MPI_SUM = 1, MPI_MIN, MPI_MAX };

namespace mpishim {

struct Message { int tag; std::vector<char> data; };
struct Pending { void* buf; size_t nbytes; int source; int tag; bool done; };

struct State {
  int rank, size;
  std::vector<int> fd;                       // fd[peer]
  std::vector<std::deque<Message> > inbox;   // per source, arrival order
  std::vector<std::vector<char> > partial;   // per source, bytes of an incomplete frame
  std::vector<Pending> reqs;
  State() : rank(0), size(1) {}
};

inline State& st() { static State s; return s; }

inline size_t type_size(MPI_Datatype t)
{
  switch (t) {
    case MPI_CHAR: case MPI_BYTE: return 1;
    case MPI_INT: case MPI_UNSIGNED: case MPI_FLOAT: return 4;
    case MPI_LONG: case MPI_LONG_LONG: case MPI_UNSIGNED_LONG: case MPI_DOUBLE: case MPI_COMPLEX: return 8;
    case MPI_DOUBLE_COMPLEX: return 16;
  }
  std::fprintf(stderr, "mpishim: unknown datatype %d\n", t); std::abort();
}

// Pull whatever is readable from peer `src` into its inbox (non-blocking).
inline void drain(int src)
{
  State& s = st();
  char tmp[1 << 16];
  for (;;) {
    ssize_t n = ::read(s.fd[src], tmp, sizeof tmp);
    if (n > 0) { s.partial[src].insert(s.partial[src].end(), tmp, tmp + n); continue; }
    if (n == 0) break;
    if (errno == EAGAIN || errno == EWOULDBLOCK) break;
    if (errno == EINTR) continue;
    std::perror("mpishim read"); std::abort();
  }
  std::vector<char>& p = s.partial[src];
  size_t off = 0;
  while (p.size() - off >= 2 * sizeof(int)) {
    int hdr[2];
    std::memcpy(hdr, &p[off], sizeof hdr);
    size_t need = 2 * sizeof(int) + (size_t)hdr[1];
    if (p.size() - off < need) break;
    Message m; m.tag = hdr[0];
    m.data.assign(p.begin() + off + 2 * sizeof(int), p.begin() + off + need);
    s.inbox[src].push_back(m);
    off += need;
  }
  if (off) p.erase(p.begin(), p.begin() + off);
}

inline void drain_all() { State& s = st(); for (int i = 0; i < s.size; ++i) if (i != s.rank) drain(i); }

inline void wait_io(int write_peer)
{
  State& s = st();
  std::vector<pollfd> pf;
  for (int i = 0; i < s.size; ++i) {
    if (i == s.rank) continue;
    pollfd p; p.fd = s.fd[i]; p.events = POLLIN; p.revents = 0;
    if (i == write_peer) p.events |= POLLOUT;
    pf.push_back(p);
  }
  ::poll(pf.empty() ? 0 : &pf[0], pf.size(), 100);
}

inline void send_bytes(int dest, int tag, const void* buf, size_t nbytes)
{
  State& s = st();
  std::vector<char> frame(2 * sizeof(int) + nbytes);
  int hdr[2] = { tag, (int)nbytes };
  std::memcpy(&frame[0], hdr, sizeof hdr);
  if (nbytes) std::memcpy(&frame[2 * sizeof(int)], buf, nbytes);
  size_t off = 0;
  while (off < frame.size()) {
    ssize_t n = ::write(s.fd[dest], &frame[off], frame.size() - off);
    if (n > 0) { off += n; continue; }
    if (n < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) {
      // keep consuming what others send us, otherwise two blocking senders deadlock
      drain_all();
      wait_io(dest);
      continue;
    }
    if (n < 0 && errno == EINTR) continue;
    std::perror("mpishim write"); std::abort();
  }
}

inline bool try_match(Pending& r, MPI_Status* status)
{
  State& s = st();
  int lo = r.source == MPI_ANY_SOURCE ? 0 : r.source;
  int hi = r.source == MPI_ANY_SOURCE ? s.size - 1 : r.source;
  for (int src = lo; src <= hi; ++src) {
    if (src == s.rank) continue;
    std::deque<Message>& q = s.inbox[src];
    for (size_t i = 0; i < q.size(); ++i) {
      if (r.tag != MPI_ANY_TAG && q[i].tag != r.tag) continue;
      if (q[i].data.size() > r.nbytes) { std::fprintf(stderr, "mpishim: message truncated\n"); std::abort(); }
      if (!q[i].data.empty()) std::memcpy(r.buf, &q[i].data[0], q[i].data.size());
      if (status) { status->MPI_SOURCE = src; status->MPI_TAG = q[i].tag; status->MPI_ERROR = MPI_SUCCESS; }
      q.erase(q.begin() + i);
      r.done = true;
      return true;
    }
  }
  return false;
}

inline void recv_bytes(int source, int tag, void* buf, size_t nbytes, MPI_Status* status)
{
  Pending r = { buf, nbytes, source, tag, false };
  for (;;) {
    drain_all();
    if (try_match(r, status)) return;
    wait_io(-1);
  }
}

template <typename T>
inline void combine(T* acc, const T* in, int count, MPI_Op op)
{
  for (int i = 0; i < count; ++i) {
    if (op == MPI_SUM) acc[i] = acc[i] + in[i];
    else if (op == MPI_MIN) acc[i] = in[i] < acc[i] ? in[i] : acc[i];
    else acc[i] = in[i] > acc[i] ? in[i] : acc[i];
  }
}

inline void combine_any(void* acc, const void* in, int count, MPI_Datatype t, MPI_Op op)
{
  switch (t) {
    case MPI_INT: combine((int*)acc, (const int*)in, count, op); break;
    case MPI_UNSIGNED: combine((unsigned*)acc, (const unsigned*)in, count, op); break;
    case MPI_LONG: combine((long*)acc, (const long*)in, count, op); break;
    case MPI_LONG_LONG: combine((long long*)acc, (const long long*)in, count, op); break;
    case MPI_UNSIGNED_LONG: combine((unsigned long*)acc, (const unsigned long*)in, count, op); break;
    case MPI_FLOAT: combine((float*)acc, (const float*)in, count, op); break;
    case MPI_DOUBLE: combine((double*)acc, (const double*)in, count, op); break;
    default: std::fprintf(stderr, "mpishim: reduce on unsupported type\n"); std::abort();
  }
}

const int TAG_COLL = 0x7f000000;

} // namespace mpishim

inline int MPI_Init(int*, char***)
{
  mpishim::State& s = mpishim::st();
  const char* r = std::getenv("MPISHIM_RANK");
  const char* n = std::getenv("MPISHIM_SIZE");
  const char* f = std::getenv("MPISHIM_FDS");
  if (!r || !n) { s.rank = 0; s.size = 1; }
  else { s.rank = std::atoi(r); s.size = std::atoi(n); }
  s.fd.assign(s.size, -1);
  s.inbox.resize(s.size);
  s.partial.resize(s.size);
  if (s.size > 1) {
    std::string fds = f ? f : "";
    size_t pos = 0;
    for (int i = 0; i < s.size; ++i) {
      size_t c = fds.find(',', pos);
      s.fd[i] = std::atoi(fds.substr(pos, c == std::string::npos ? c : c - pos).c_str());
      pos = c == std::string::npos ? fds.size() : c + 1;
      if (i != s.rank) {
        int fl = ::fcntl(s.fd[i], F_GETFL, 0);
        ::fcntl(s.fd[i], F_SETFL, fl | O_NONBLOCK);
      }
    }
  }
  return MPI_SUCCESS;
}
inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = mpishim::st().rank; return MPI_SUCCESS; }
inline int MPI_Comm_size(MPI_Comm, int* n) { *n = mpishim::st().size; return MPI_SUCCESS; }
inline int MPI_Abort(MPI_Comm, int code) { std::fprintf(stderr, "mpishim: MPI_Abort(%d)\n", code); std::_Exit(code ? code : 1); }
inline double MPI_Wtime() { timeval tv; gettimeofday(&tv, 0); return tv.tv_sec + 1e-6 * tv.tv_usec; }
inline int MPI_Pcontrol(int, ...) { return MPI_SUCCESS; }

inline int MPI_Send(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm)
{
  mpishim::send_bytes(dest, tag, buf, (size_t)count * mpishim::type_size(t));
  return MPI_SUCCESS;
}
inline int MPI_Recv(void* buf, int count, MPI_Datatype t, int source, int tag, MPI_Comm, MPI_Status* status)
{
  mpishim::recv_bytes(source, tag, buf, (size_t)count * mpishim::type_size(t), status);
  return MPI_SUCCESS;
}
inline int MPI_Irecv(void* buf, int count, MPI_Datatype t, int source, int tag, MPI_Comm, MPI_Request* req)
{
  mpishim::State& s = mpishim::st();
  mpishim::Pending p = { buf, (size_t)count * mpishim::type_size(t), source, tag, false };
  s.reqs.push_back(p);
  *req = (int)s.reqs.size() - 1;
  return MPI_SUCCESS;
}
inline int MPI_Wait(MPI_Request* req, MPI_Status* status)
{
  mpishim::State& s = mpishim::st();
  mpishim::Pending& p = s.reqs[*req];
  while (!p.done) {
    mpishim::drain_all();
    if (mpishim::try_match(p, status)) break;
    mpishim::wait_io(-1);
  }
  return MPI_SUCCESS;
}

inline int MPI_Bcast(void* buf, int count, MPI_Datatype t, int root, MPI_Comm)
{
  mpishim::State& s = mpishim::st();
  size_t nb = (size_t)count * mpishim::type_size(t);
  if (s.rank == root) { for (int i = 0; i < s.size; ++i) if (i != root) mpishim::send_bytes(i, mpishim::TAG_COLL + 1, buf, nb); }
  else mpishim::recv_bytes(root, mpishim::TAG_COLL + 1, buf, nb, 0);
  return MPI_SUCCESS;
}
inline int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm)
{
  mpishim::State& s = mpishim::st();
  size_t nb = (size_t)count * mpishim::type_size(t);
  if (s.rank != root) { mpishim::send_bytes(root, mpishim::TAG_COLL + 2, sendbuf, nb); return MPI_SUCCESS; }
  // rank-ordered association: ((v0 + v1) + v2) + ...
  std::vector<std::vector<char> > contrib(s.size, std::vector<char>(nb));
  for (int i = 0; i < s.size; ++i) {
    if (i == root) { if (nb) std::memcpy(&contrib[i][0], sendbuf, nb); }
    else mpishim::recv_bytes(i, mpishim::TAG_COLL + 2, nb ? &contrib[i][0] : 0, nb, 0);
  }
  std::vector<char> acc = contrib[0];
  for (int i = 1; i < s.size; ++i) mpishim::combine_any(nb ? &acc[0] : 0, nb ? &contrib[i][0] : 0, count, t, op);
  if (nb) std::memcpy(recvbuf, &acc[0], nb);
  return MPI_SUCCESS;
}
inline int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{
  std::vector<char> tmp((size_t)count * mpishim::type_size(t));
  if (!tmp.empty()) std::memcpy(&tmp[0], sendbuf, tmp.size());   // sendbuf may alias recvbuf
  MPI_Reduce(tmp.empty() ? 0 : &tmp[0], recvbuf, count, t, op, 0, c);
  return MPI_Bcast(recvbuf, count, t, 0, c);
}
inline int MPI_Allgather(const void* sendbuf, int scount, MPI_Datatype st_, void* recvbuf, int rcount, MPI_Datatype rt, MPI_Comm c)
{
  mpishim::State& s = mpishim::st();
  size_t nb = (size_t)scount * mpishim::type_size(st_);
  char* out = (char*)recvbuf;
  if (s.rank == 0) {
    std::memcpy(out, sendbuf, nb);
    for (int i = 1; i < s.size; ++i) mpishim::recv_bytes(i, mpishim::TAG_COLL + 3, out + i * nb, nb, 0);
  } else mpishim::send_bytes(0, mpishim::TAG_COLL + 3, sendbuf, nb);
  return MPI_Bcast(recvbuf, rcount * s.size, rt, 0, c);
}
inline int MPI_Barrier(MPI_Comm c) { int a = 0, b = 0; return MPI_Allreduce(&a, &b, 1, MPI_INT, MPI_SUM, c); }
inline int MPI_Finalize() { MPI_Barrier(MPI_COMM_WORLD); return MPI_SUCCESS; }

#endif
