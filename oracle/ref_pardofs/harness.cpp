// harness.cpp -- runs the REFERENCE's ParallelDofs (linalg/paralleldofs.cpp:20-108, linalg/paralleldofs.hpp:213-334, compiled
// from the sources where they lie with -DPARALLEL -DNG_MPI_WRAPPER) on N ranks inside one process, to pin the distributed
// index maps of the library against reference output.  The image has no MPI: netgen's MPI layer is a table of function
// pointers (netgen/libsrc/core/ng_mpi.hpp, ng_mpi_generated_declarations.hpp); this file defines the ones the class uses
// with "one thread = one rank" semantics (mailboxes in process memory).  Test infrastructure only (see oracle/Makefile).
//
//   ref_pardofs <in.bin> <out.bin>
// in : int32 nranks, then per rank: int32 ndof, int32 first[ndof+1], int32 dist_procs[first[ndof]], double data[ndof]
// out: per rank: int64 global_ndof, int32 ex_first[nranks+1], int32 ex_dofs[...], uint8 ismaster[ndof],
//                double allreduced[ndof] (AllReduceDofData(data, SUM), linalg/jacobi.cpp:60-61),
//                double scattered[ndof] (ReduceDofData(SUM) then ScatterDofData)
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include <la.hpp>          // brings ngstd + paralleldofs.hpp in PARALLEL form

using namespace ngcore;

// ---------------------------------------------------------------------------------------------------------------------
// threads-as-ranks message passing
// ---------------------------------------------------------------------------------------------------------------------
static int g_nranks = 1;
static thread_local int t_rank = 0;

struct Message { int src, tag; std::vector<char> bytes; };
static std::mutex g_mu;
static std::condition_variable g_cv;
static std::vector<std::deque<Message>> g_box;       // per destination rank

struct IndexedType { size_t elem_bytes; std::vector<int> displs; };     // MPI_Type_indexed with block length 1
static std::map<uintptr_t, size_t> g_basic = {};                          // handle -> bytes
static std::map<uintptr_t, IndexedType> g_indexed;
static std::map<uintptr_t, size_t> g_contig;
static uintptr_t g_next_type = 1000;
static const uintptr_t T_DOUBLE = 1, T_INT = 2, T_SIZET = 3, T_CHAR = 4, T_COMPLEX = 5, T_SHORT = 6, T_BOOL = 7;

static size_t type_bytes (NG_MPI_Datatype t)
{
  switch (t.value) { case T_DOUBLE: return 8; case T_INT: return 4; case T_SIZET: return 8; case T_CHAR: return 1; case T_COMPLEX: return 16;
                     case T_SHORT: return 2; case T_BOOL: return 1; }
  std::lock_guard<std::mutex> l(g_mu);
  if (g_contig.count (t.value)) return g_contig[t.value];
  throw std::runtime_error ("shim: unknown datatype");
}

struct PendingRecv { void * buf; size_t bytes; int src, tag; };
static std::mutex g_req_mu;
static std::map<uintptr_t, PendingRecv> g_recvs;
static uintptr_t g_next_req = 1;

static void deliver (int dest, int src, int tag, const void * buf, size_t bytes)
{
  Message m { src, tag, std::vector<char> ((const char*)buf, (const char*)buf + bytes) };
  { std::lock_guard<std::mutex> l(g_mu); g_box[dest].push_back (std::move (m)); }
  g_cv.notify_all ();
}
static void receive (int me, int src, int tag, void * buf, size_t bytes)
{
  std::unique_lock<std::mutex> l(g_mu);
  for (;;)
    {
      auto & q = g_box[me];
      for (auto it = q.begin(); it != q.end(); ++it)
        if (it->src == src && it->tag == tag)
          {
            if (it->bytes.size() != bytes) throw std::runtime_error ("shim: message size mismatch");
            memcpy (buf, it->bytes.data(), bytes);
            q.erase (it);
            return;
          }
      g_cv.wait (l);
    }
}

// barrier + all-reduce through a shared accumulator
static std::mutex g_coll_mu;
static std::condition_variable g_coll_cv;
static int g_coll_count = 0, g_coll_gen = 0;
static std::vector<char> g_coll_buf;
template <typename F, typename G> static void collective (F contribute, G fetch)
{
  std::unique_lock<std::mutex> l(g_coll_mu);
  int gen = g_coll_gen;
  contribute ();
  if (++g_coll_count == g_nranks) { g_coll_count = 0; g_coll_gen++; g_coll_cv.notify_all (); }
  else g_coll_cv.wait (l, [&] { return g_coll_gen != gen; });
  fetch ();
  // second phase so that the buffer is not reused before everybody fetched
  gen = g_coll_gen;
  if (++g_coll_count == g_nranks) { g_coll_count = 0; g_coll_gen++; g_coll_buf.clear (); g_coll_cv.notify_all (); }
  else g_coll_cv.wait (l, [&] { return g_coll_gen != gen; });
}

namespace ngcore
{
  // the entries of the function table that ParallelDofs and NgMPI_Comm touch (all others stay undefined: a link error
  // would name any further one)
  int (*NG_MPI_Comm_rank)(NG_MPI_Comm, int*) = [] (NG_MPI_Comm, int * r) { *r = t_rank; return 0; };
  int (*NG_MPI_Comm_size)(NG_MPI_Comm, int* s) = [] (NG_MPI_Comm, int * s) { *s = g_nranks; return 0; };
  int (*NG_MPI_Comm_free)(NG_MPI_Comm*) = [] (NG_MPI_Comm*) { return 0; };
  int (*NG_MPI_Barrier)(NG_MPI_Comm) = [] (NG_MPI_Comm) { collective ([] {}, [] {}); return 0; };
  int (*NG_MPI_Type_contiguous)(int, NG_MPI_Datatype, NG_MPI_Datatype*) = [] (int n, NG_MPI_Datatype t, NG_MPI_Datatype * out)
  { size_t b = type_bytes (t) * n; std::lock_guard<std::mutex> l(g_mu); out->value = g_next_type++; g_contig[out->value] = b; return 0; };
  int (*NG_MPI_Type_indexed)(int, int*, int*, NG_MPI_Datatype, NG_MPI_Datatype*) = [] (int n, int * bl, int * displ, NG_MPI_Datatype t, NG_MPI_Datatype * out)
  {
    size_t b = type_bytes (t);
    for (int i = 0; i < n; i++) if (bl[i] != 1) throw std::runtime_error ("shim: block length != 1");
    std::lock_guard<std::mutex> l(g_mu);
    out->value = g_next_type++;
    g_indexed[out->value] = IndexedType { b, std::vector<int> (displ, displ + n) };
    return 0;
  };
  int (*NG_MPI_Type_commit)(NG_MPI_Datatype*) = [] (NG_MPI_Datatype*) { return 0; };
  int (*NG_MPI_Type_free)(NG_MPI_Datatype*) = [] (NG_MPI_Datatype*) { return 0; };
  int (*NG_MPI_Isend)(void*, int, NG_MPI_Datatype, int, int, NG_MPI_Comm, NG_MPI_Request*) =
    [] (void * buf, int count, NG_MPI_Datatype t, int dest, int tag, NG_MPI_Comm, NG_MPI_Request * req)
  { deliver (dest, t_rank, tag, buf, type_bytes (t) * count); req->value = 0; return 0; };      // buffered: complete at once
  int (*NG_MPI_Irecv)(void*, int, NG_MPI_Datatype, int, int, NG_MPI_Comm, NG_MPI_Request*) =
    [] (void * buf, int count, NG_MPI_Datatype t, int src, int tag, NG_MPI_Comm, NG_MPI_Request * req)
  {
    std::lock_guard<std::mutex> l(g_req_mu);
    req->value = g_next_req++;
    g_recvs[req->value] = PendingRecv { buf, type_bytes (t) * count, src, tag };
    return 0;
  };
  static void wait_one (NG_MPI_Request * req)
  {
    if (req->value == 0) return;
    PendingRecv p;
    { std::lock_guard<std::mutex> l(g_req_mu); p = g_recvs[req->value]; g_recvs.erase (req->value); }
    receive (t_rank, p.src, p.tag, p.buf, p.bytes);
    req->value = 0;
  }
  int (*NG_MPI_Wait)(NG_MPI_Request*, NG_MPI_Status*) = [] (NG_MPI_Request * r, NG_MPI_Status*) { wait_one (r); return 0; };
  int (*NG_MPI_Waitall)(int, NG_MPI_Request*, NG_MPI_Status*) = [] (int n, NG_MPI_Request * r, NG_MPI_Status*)
  { for (int i = 0; i < n; i++) wait_one (r + i); return 0; };
  int (*NG_MPI_Waitany)(int, NG_MPI_Request*, int*, NG_MPI_Status*) = [] (int n, NG_MPI_Request * r, int * idx, NG_MPI_Status*)
  { for (int i = 0; i < n; i++) if (r[i].value) { wait_one (r + i); *idx = i; return 0; } *idx = -1; return 0; };
  int (*NG_MPI_Reduce_local)(void*, void*, int, NG_MPI_Datatype, NG_MPI_Op) = [] (void * in, void * inout, int n, NG_MPI_Datatype t, NG_MPI_Op op)
  {
    if (t.value != T_DOUBLE || op.value != 1) throw std::runtime_error ("shim: Reduce_local only for double SUM");
    for (int i = 0; i < n; i++) ((double*)inout)[i] = ((double*)in)[i] + ((double*)inout)[i];      // MPI: inout = in op inout
    return 0;
  };
  int (*NG_MPI_Allreduce)(void*, void*, int, NG_MPI_Datatype, NG_MPI_Op, NG_MPI_Comm) = [] (void * in, void * out, int n, NG_MPI_Datatype t, NG_MPI_Op op, NG_MPI_Comm)
  {
    if (op.value != 1 || n != 1 || (t.value != T_SIZET && t.value != T_INT && t.value != T_DOUBLE)) throw std::runtime_error ("shim: Allreduce subset");
    size_t b = type_bytes (t);
    // rank order does not matter for integer sums; doubles are not reduced this way by the code under test
    collective ([&] {
                  if (g_coll_buf.empty ()) g_coll_buf.assign (16, 0);
                  if (t.value == T_SIZET) *(size_t*)g_coll_buf.data() += *(size_t*)in;
                  else if (t.value == T_INT) *(int*)g_coll_buf.data() += *(int*)in;
                  else *(double*)g_coll_buf.data() += *(double*)in;
                },
                [&] { memcpy (out, g_coll_buf.data(), b); });
    return 0;
  };
  int (*NG_MPI_Initialized)(int*) = [] (int * flag) { *flag = 1; return 0; };
  int (*NG_MPI_Allgather)(void*, int, NG_MPI_Datatype, void*, int, NG_MPI_Datatype, NG_MPI_Comm) =
    [] (void * in, int n, NG_MPI_Datatype t, void * out, int, NG_MPI_Datatype, NG_MPI_Comm)
  {
    size_t b = type_bytes (t) * n;
    collective ([&] { if (g_coll_buf.size () != b * g_nranks) g_coll_buf.assign (b * g_nranks, 0); memcpy (g_coll_buf.data () + b * t_rank, in, b); },
                [&] { memcpy (out, g_coll_buf.data (), b * g_nranks); });
    return 0;
  };
  NG_MPI_Datatype NG_MPI_DOUBLE = T_DOUBLE, NG_MPI_INT = T_INT, NG_MPI_UINT64_T = T_SIZET, NG_MPI_CHAR = T_CHAR, NG_MPI_CXX_DOUBLE_COMPLEX = T_COMPLEX,
    NG_MPI_SHORT = T_SHORT, NG_MPI_C_BOOL = T_BOOL, NG_MPI_DATATYPE_NULL = 0, NG_MPI_FLOAT = 8;
  NG_MPI_Op NG_MPI_SUM = 1, NG_MPI_MAX = 2, NG_MPI_MIN = 3, NG_MPI_LOR = 4;
  NG_MPI_Comm NG_MPI_COMM_WORLD = 1, NG_MPI_COMM_NULL = 0;
  void* NG_MPI_IN_PLACE = (void*)-1;
  NG_MPI_Status* NG_MPI_STATUS_IGNORE = nullptr;
  NG_MPI_Status* NG_MPI_STATUSES_IGNORE = nullptr;
  NG_MPI_Request NG_MPI_REQUEST_NULL = 0;
}

// ---------------------------------------------------------------------------------------------------------------------
struct RankIn { int ndof; std::vector<int> first, dp; std::vector<double> data; };
struct RankOut { long long global_ndof; std::vector<int> ex_first, ex_dofs; std::vector<unsigned char> master; std::vector<double> allred, scat; };

int main (int argc, char ** argv)
{
  if (argc < 3) { fprintf (stderr, "usage: ref_pardofs in.bin out.bin\n"); return 2; }
  FILE * f = fopen (argv[1], "rb");
  if (!f) { perror ("in"); return 2; }
  int np = 0;
  if (fread (&np, 4, 1, f) != 1) return 2;
  std::vector<RankIn> in (np);
  for (auto & r : in)
    {
      if (fread (&r.ndof, 4, 1, f) != 1) return 2;
      r.first.resize (r.ndof + 1);
      if (fread (r.first.data(), 4, r.ndof + 1, f) != size_t(r.ndof + 1)) return 2;
      r.dp.resize (r.first[r.ndof]);
      if (!r.dp.empty () && fread (r.dp.data(), 4, r.dp.size(), f) != r.dp.size()) return 2;
      r.data.resize (r.ndof);
      if (fread (r.data.data(), 8, r.ndof, f) != size_t(r.ndof)) return 2;
    }
  fclose (f);
  g_nranks = np;
  g_box.resize (np);
  std::vector<RankOut> out (np);
  std::vector<std::string> errors (np);
  std::vector<std::thread> th;
  for (int rank = 0; rank < np; rank++)
    th.emplace_back ([&, rank] {
      t_rank = rank;
      try
        {
          const RankIn & r = in[rank];
          Array<int> cnt (r.ndof);
          for (int i = 0; i < r.ndof; i++) cnt[i] = r.first[i + 1] - r.first[i];
          Table<int> dist_procs (cnt);
          for (int i = 0; i < r.ndof; i++)
            for (int k = 0; k < cnt[i]; k++) dist_procs[i][k] = r.dp[r.first[i] + k];
          NgMPI_Comm comm (NG_MPI_COMM_WORLD, false);
          ngla::ParallelDofs pd (comm, std::move (dist_procs), 1, false);          // the reference constructor
          RankOut & o = out[rank];
          o.global_ndof = pd.GetNDofGlobal ();
          o.ex_first.assign (np + 1, 0);
          for (int p = 0; p < np; p++)
            {
              auto ex = pd.GetExchangeDofs (p);
              o.ex_first[p + 1] = o.ex_first[p] + int (ex.Size ());
              for (auto d : ex) o.ex_dofs.push_back (d);
            }
          o.master.resize (r.ndof);
          for (int i = 0; i < r.ndof; i++) o.master[i] = pd.IsMasterDof (i) ? 1 : 0;
          o.allred = r.data;
          pd.AllReduceDofData (FlatArray<double> (r.ndof, o.allred.data ()), NG_MPI_SUM);
          o.scat = r.data;
          pd.ReduceDofData (FlatArray<double> (r.ndof, o.scat.data ()), NG_MPI_SUM);
          pd.ScatterDofData (FlatArray<double> (r.ndof, o.scat.data ()));
        }
      catch (const std::exception & e) { errors[rank] = e.what (); }
    });
  for (auto & t : th) t.join ();
  for (int rank = 0; rank < np; rank++)
    if (!errors[rank].empty ()) { fprintf (stderr, "rank %d: %s\n", rank, errors[rank].c_str ()); return 1; }
  FILE * g = fopen (argv[2], "wb");
  if (!g) { perror ("out"); return 2; }
  for (auto & o : out)
    {
      fwrite (&o.global_ndof, 8, 1, g);
      fwrite (o.ex_first.data (), 4, o.ex_first.size (), g);
      if (!o.ex_dofs.empty ()) fwrite (o.ex_dofs.data (), 4, o.ex_dofs.size (), g);
      fwrite (o.master.data (), 1, o.master.size (), g);
      fwrite (o.allred.data (), 8, o.allred.size (), g);
      fwrite (o.scat.data (), 8, o.scat.size (), g);
    }
  fclose (g);
  return 0;
}
