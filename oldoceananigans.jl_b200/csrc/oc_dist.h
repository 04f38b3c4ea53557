// oc_dist.h — slab (y) domain decomposition over R GPUs: halo exchange and the transposed distributed FFT.
//
// Replaces, for Partition(1, R) on a RectilinearGrid that is Periodic in y (src/DistributedComputations):
//   halo_communication.jl:87-333 + communication_buffers.jl:30-135  fill_halo_regions! with MPI Isend/Irecv of packed y-halos
//   distributed_transpose.jl:25-191 + transposable_field.jl:49-105   pack -> Alltoallv -> unpack transposes
//   distributed_fft_based_poisson_solver.jl:92-188                   FFT(z,x local) -> transpose -> FFT(y) -> divide -> back
// One process per GPU; NCCL (dlopen'ed: libnccl.so.2 — no link-time dependency for single-GPU users) carries both patterns
// as grouped ncclSend/ncclRecv.  A host-callback transport exists for the TEST-ONLY host simulation (world_size-2 gloo tests).
//
// Layouts.  Local spectral buffer (as on one GPU): complex (nxc, Ny_l, Nz), x fastest — the z-chunk destined to rank d is
// CONTIGUOUS, so the forward all-to-all needs no pack.  Received blocks land in `stage` as [s][zl][yl][x]; TransposeUnpack
// turns them into T = (y fastest, x, zl) through a shared-memory tile so that both sides are coalesced and the y-FFT is a
// contiguous batched 1-D transform.  The way back mirrors it.
#pragma once
#include "oc_halo.h"
#include "oc_poisson.h"

#ifndef OC_HOSTSIM
#include <cufft.h>
#include <dlfcn.h>
#endif

namespace oc {

struct Msg {
    int send_peer, recv_peer, tag;     // messages between one pair of ranks are matched in posting order (tag: host transport)
    void* send;
    size_t send_bytes;
    void* recv;
    size_t recv_bytes;
};

// oc_exchange_fn of include/oceananigans_b200.h: perform all transfers, return 0 on success
typedef int (*HostExchangeFn)(void*, int, const int*, const int*, const int*, void* const*, const size_t*, void* const*, const size_t*);

struct Transport {
    virtual ~Transport() {}
    virtual std::string exchange(const std::vector<Msg>& msgs, Stream stream) = 0;
    // Peer memory (NVLink / NVSwitch): map every rank's allocation `local` (a cudaMalloc base, the same call on every rank, in the same
    // order) into this process; peers[r] is rank r's buffer (peers[rank] = local).  Returns "" on success; transports without peer
    // memory return a reason and the caller keeps the send / receive path.
    virtual std::string map_peers(void* local, std::vector<void*>& peers, Stream stream) { (void)local; (void)peers; (void)stream; return "no peer memory in this transport"; }
    // every rank's work enqueued before the barrier is complete (and its peer writes visible) before any rank's work enqueued after it starts
    virtual std::string barrier(Stream stream) { (void)stream; return "no barrier in this transport"; }
    virtual int agree(bool failed, Stream stream) { (void)stream; return failed ? 1 : 0; }
};

struct HostTransport : Transport {
    HostExchangeFn fn;
    void* user;
    int rank, nranks;
    // TEST-ONLY (host simulation, OC_HOSTSIM_THREADS=1): the ranks are threads of ONE process, so a peer's buffer is simply its pointer
    // and the peer-memory transposes (TransposePutKernel) can be exercised without a GPU; the collectives go through the callback
    bool same_process;
    HostTransport(HostExchangeFn f, void* u, int r, int n, bool threads) : fn(f), user(u), rank(r), nranks(n), same_process(threads) {}
    std::string all_gather(const void* mine, void* all, size_t bytes, int tag) {
        memcpy((char*)all + bytes * rank, mine, bytes);
        std::vector<Msg> msgs;
        for (int d = 1; d < nranks; ++d) {
            const int to = (rank + d) % nranks, from = (rank + nranks - d) % nranks;
            msgs.push_back(Msg{to, from, tag + d, (char*)all + bytes * rank, bytes, (char*)all + bytes * from, bytes});
        }
        return exchange(msgs, Stream());
    }
    std::string map_peers(void* local, std::vector<void*>& peers, Stream) override {
        if (!same_process) return "no peer memory in this transport";
        peers.assign(nranks, nullptr);
        return all_gather(&local, peers.data(), sizeof(void*), 900);
    }
    std::string barrier(Stream) override {
        if (!same_process) return "no barrier in this transport";
        char mine = 0;
        std::vector<char> all(nranks, 0);
        return all_gather(&mine, all.data(), 1, 940);
    }
    int agree(bool failed, Stream) override {
        if (!same_process) return failed ? 1 : 0;
        int mine = failed ? 1 : 0, sum = 0;
        std::vector<int> all(nranks, 0);
        if (!all_gather(&mine, all.data(), sizeof(int), 980).empty()) return nranks;
        for (int v : all) sum += v;
        return sum;
    }
    std::string exchange(const std::vector<Msg>& msgs, Stream) override {
        std::vector<int> sp, rp, tg;
        std::vector<void*> sptr, rptr;
        std::vector<size_t> sb, rb;
        for (const Msg& m : msgs) {
            sp.push_back(m.send_peer); rp.push_back(m.recv_peer); tg.push_back(m.tag);
            sptr.push_back(m.send); sb.push_back(m.send_bytes); rptr.push_back(m.recv); rb.push_back(m.recv_bytes);
        }
        int rc = fn(user, (int)msgs.size(), sp.data(), rp.data(), tg.data(), sptr.data(), sb.data(), rptr.data(), rb.data());
        return rc == 0 ? "" : "host exchange callback failed with code " + std::to_string(rc);
    }
};

#ifndef OC_HOSTSIM
// The handful of NCCL entry points, resolved at run time.
struct NcclApi {
    typedef struct ncclComm* comm_t;
    struct UniqueId { char internal[128]; };
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(comm_t*, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string load() {
        if (GetUniqueId) return "";
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return std::string("cannot load libnccl.so.2: ") + dlerror();
#define OC_NCCL_SYM(field, name) *(void**)(&field) = dlsym(h, name); if (!field) return std::string("libnccl lacks ") + name;
        OC_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        OC_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        OC_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        OC_NCCL_SYM(Send, "ncclSend")
        OC_NCCL_SYM(Recv, "ncclRecv")
        OC_NCCL_SYM(AllReduce, "ncclAllReduce")
        OC_NCCL_SYM(GroupStart, "ncclGroupStart")
        OC_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        OC_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef OC_NCCL_SYM
        return "";
    }
};
inline NcclApi& nccl_api() { static NcclApi api; return api; }

struct NcclTransport : Transport {
    NcclApi::comm_t comm = nullptr;
    int rank_ = 0, nranks_ = 1;
    int* flag_ = nullptr;                  // device word of the barrier's all-reduce
    std::vector<void*> mapped_;            // peer allocations opened with cudaIpcOpenMemHandle
    std::string init(int rank, int nranks, const void* id128) {
        std::string e = nccl_api().load();
        if (!e.empty()) return e;
        NcclApi::UniqueId id;
        memcpy(id.internal, id128, 128);
        rank_ = rank; nranks_ = nranks;
        int rc = nccl_api().CommInitRank(&comm, nranks, id, rank);
        if (rc != 0) return std::string("ncclCommInitRank: ") + nccl_api().GetErrorString(rc);
        if (cudaMalloc((void**)&flag_, 2 * sizeof(int)) != cudaSuccess || cudaMemset(flag_, 0, 2 * sizeof(int)) != cudaSuccess) return "cudaMalloc(barrier word) failed";
        return "";
    }
    ~NcclTransport() override {
        for (void* p : mapped_) cudaIpcCloseMemHandle(p);
        if (flag_) cudaFree(flag_);
        if (comm) nccl_api().CommDestroy(comm);
    }
    // CUDA IPC: one process per GPU, so a peer's cudaMalloc'ed buffer is reached through its IPC handle; the 64-byte handles travel
    // over the communicator (grouped send / receive through a small device buffer)
    // (Collective: a rank whose own steps fail still takes part in the exchange, so that nobody hangs; agree() then settles the outcome.)
    std::string map_peers(void* local, std::vector<void*>& peers, Stream stream) override {
        peers.assign(nranks_, nullptr);
        std::string err;
        cudaIpcMemHandle_t mine;
        memset(&mine, 0, sizeof(mine));
        if (cudaIpcGetMemHandle(&mine, local) != cudaSuccess) { cudaGetLastError(); err = "cudaIpcGetMemHandle failed"; }
        const size_t hb = sizeof(cudaIpcMemHandle_t);
        char* dbuf = nullptr;
        if (cudaMalloc((void**)&dbuf, hb * nranks_) != cudaSuccess) return "cudaMalloc(IPC handles) failed";
        std::vector<cudaIpcMemHandle_t> all(nranks_);
        if (cudaMemcpyAsync(dbuf + hb * rank_, &mine, hb, cudaMemcpyHostToDevice, stream) != cudaSuccess && err.empty()) err = "cudaMemcpy(IPC handle) failed";
        {
            std::vector<Msg> msgs;
            for (int d = 1; d < nranks_; ++d) {
                const int to = (rank_ + d) % nranks_, from = (rank_ + nranks_ - d) % nranks_;
                msgs.push_back(Msg{to, from, 900 + d, dbuf + hb * rank_, hb, dbuf + hb * from, hb});
            }
            std::string e = exchange(msgs, stream);
            if (err.empty()) err = e;
        }
        if ((cudaMemcpyAsync(all.data(), dbuf, hb * nranks_, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
             cudaStreamSynchronize(stream) != cudaSuccess) && err.empty()) err = "cudaMemcpy(IPC handles) failed";
        cudaFree(dbuf);
        if (!err.empty()) return err;
        for (int r = 0; r < nranks_; ++r) {
            if (r == rank_) { peers[r] = local; continue; }
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                return "cudaIpcOpenMemHandle(rank " + std::to_string(r) + ") failed";
            }
            mapped_.push_back(p);
            peers[r] = p;
        }
        return "";
    }
    // number of ranks that report a failure (collective; synchronises the stream)
    int agree(bool failed, Stream stream) override {
        int v[2] = {failed ? 1 : 0, 0};
        if (cudaMemcpyAsync(flag_, v, sizeof(int), cudaMemcpyHostToDevice, stream) != cudaSuccess) return nranks_;
        if (nccl_api().AllReduce(flag_, flag_ + 1, 1, /*ncclInt32*/ 2, /*ncclSum*/ 0, comm, stream) != 0) return nranks_;
        if (cudaMemcpyAsync(&v[1], flag_ + 1, sizeof(int), cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) return nranks_;
        cudaMemsetAsync(flag_, 0, 2 * sizeof(int), stream);
        return v[1];
    }
    std::string barrier(Stream stream) override {
        int rc = nccl_api().AllReduce(flag_, flag_ + 1, 1, /*ncclInt32*/ 2, /*ncclSum*/ 0, comm, stream);
        return rc == 0 ? "" : std::string("NCCL all-reduce (barrier): ") + nccl_api().GetErrorString(rc);
    }
    std::string exchange(const std::vector<Msg>& msgs, Stream stream) override {
        NcclApi& n = nccl_api();
        int rc = n.GroupStart();
        for (const Msg& m : msgs) {
            if (rc == 0 && m.recv_bytes) rc = n.Recv(m.recv, m.recv_bytes, /*ncclChar*/ 0, m.recv_peer, comm, stream);
            if (rc == 0 && m.send_bytes) rc = n.Send(m.send, m.send_bytes, 0, m.send_peer, comm, stream);
        }
        int rc2 = n.GroupEnd();
        if (rc == 0) rc = rc2;
        return rc == 0 ? "" : std::string("NCCL send/recv: ") + n.GetErrorString(rc);
    }
};
#endif

// ---------------------------------------------------------------------------------------------------------
// y-halo pack / unpack: all fields, both sides, one launch.  A slab is H rows × the whole x pitch × every plane.
// buffer layout: [side][field][plane][row][x]
// ---------------------------------------------------------------------------------------------------------
template <class FT>
struct HaloPackKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    Geom<FT> g;
    int nfields, planes, rows;     // rows = halo width exchanged
    int unpack;                    // 0: interior edge rows -> buffer ; 1: buffer -> halo rows
    FT* base[HALO_MAX_FIELDS];     // allocation bases
    FT* buf;                       // [2][nfields][planes][rows][sy]
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        const long long per_side = (long long)nfields * planes * rows * g.sy;
        long long n = (long long)b.x * nt + tid;
        if (n >= 2 * per_side) return;
        const int side = (int)(n / per_side);
        long long r = n - side * per_side;
        const int x = (int)(r % g.sy); r /= g.sy;
        const int row = (int)(r % rows); r /= rows;
        const int pl = (int)(r % planes);
        const int f = (int)(r / planes);
        if (unpack && (side == 0 ? g.wlo[1] : g.whi[1])) return;       // a wall side (Bounded y, outer rank): nothing was received
        // allocation row index of interior row j is j + H[1]
        int jrow;
        if (!unpack) jrow = side == 0 ? g.H[1] + row : g.H[1] + g.N[1] - rows + row;          // low edge / high edge interior rows
        else jrow = side == 0 ? g.H[1] - rows + row : g.H[1] + g.N[1] + row;                   // low halo / high halo rows
        FT* p = base[f] + (long long)pl * g.sz + (long long)jrow * g.sy + x;
        if (!unpack) buf[n] = *p; else *p = buf[n];
    }
};

// x-halo pack / unpack (pencil decompositions): H columns of EVERY row of the parent array — wall-side boundary-condition halos
// included — and every plane.  buffer layout: [side][field][plane][row][column]
template <class FT>
struct HaloPackXKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    Geom<FT> g;
    int nfields, planes, rows, cols;   // rows = rows of the allocation (sz / sy); cols = halo width exchanged
    int unpack;                        // 0: interior edge columns -> buffer ; 1: buffer -> halo columns
    FT* base[HALO_MAX_FIELDS];         // allocation bases
    int x0[HALO_MAX_FIELDS];           // offset of interior column 0 within a row
    FT* buf;
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        const long long per_side = (long long)nfields * planes * rows * cols;
        long long n = (long long)b.x * nt + tid;
        if (n >= 2 * per_side) return;
        const int side = (int)(n / per_side);
        long long r = n - side * per_side;
        const int c = (int)(r % cols); r /= cols;
        const int row = (int)(r % rows); r /= rows;
        const int pl = (int)(r % planes);
        const int f = (int)(r / planes);
        if (unpack && (side == 0 ? g.wlo[0] : g.whi[0])) return;       // a wall side: nothing was received
        int col;
        if (!unpack) col = side == 0 ? c : g.N[0] - cols + c;          // low edge / high edge interior columns
        else col = side == 0 ? c - cols : g.N[0] + c;                  // low halo / high halo columns
        FT* p = base[f] + (long long)pl * g.sz + (long long)row * g.sy + x0[f] + col;
        if (!unpack) buf[n] = *p; else *p = buf[n];
    }
};

// ---------------------------------------------------------------------------------------------------------
// transposes between stage = [s][zl][yl][x] (x fastest) and T = [zl][x][y] (y fastest), y = s·Ny_l + yl.  32×32 tiles.
// ---------------------------------------------------------------------------------------------------------
template <class FT>
struct TransposeKernel {
    static constexpr int PHASES = 2;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr size_t SMEM = sizeof(Cplx<FT>) * 32 * 33;
    int nxc, nyl, nzl, R;
    int zl0;                   // first local z-level of this launch (grid.z levels from there): sub-chunk pipelining
    int to_T;                  // 1: stage -> T ; 0: T -> stage
    int yperm;                 // Bounded y: T holds the Makhoul-permuted line (v[n] = x[2n], v[N-1-n] = x[2n+1]) the DCT's FFT runs on
    Cplx<FT>* stage;
    Cplx<FT>* T;
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char* smem) const {
        Cplx<FT>* tile = reinterpret_cast<Cplx<FT>*>(smem);
        const int ny = nyl * R;
        const int x0 = b.x * 32, y0 = b.y * 32, zl = zl0 + b.z;
        const int tx = tid & 31, ty = tid >> 5;        // 32 × 8
        if ((PHASE == 0) == (to_T != 0)) {
            // touch the stage side (x contiguous): PHASE 0 reads it when to_T, PHASE 1 writes it when !to_T
            for (int r = ty; r < 32; r += 8) {
                const int x = x0 + tx, y = y0 + r;
                if (x < nxc && y < ny) {
                    const int s = y / nyl, yl = y - s * nyl;
                    Cplx<FT>* p = stage + ((((long long)s * nzl + zl) * nyl + yl) * nxc + x);
                    if (to_T) tile[r * 33 + tx] = *p; else *p = tile[r * 33 + tx];
                }
            }
        } else {
            // touch the T side (y contiguous)
            for (int r = ty; r < 32; r += 8) {
                const int x = x0 + r, y = y0 + tx;
                if (x < nxc && y < ny) {
                    Cplx<FT>* p = T + (((long long)zl * nxc + x) * ny + makhoul(y, ny, yperm));
                    if (to_T) *p = tile[tx * 33 + r]; else tile[tx * 33 + r] = *p;
                }
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------
// The transposes of the distributed solve as ONE kernel each over peer memory (NVLink / NVSwitch), replacing pack -> all-to-all ->
// unpack (distributed_transpose.jl:25-191: pack_buffer!, Alltoallv, unpack_buffer!): every CTA reads a 32×32 tile of the local
// buffer through shared memory and stores it — transposed — straight into the buffer of the rank that owns it.  The transfer IS the
// transpose: no staging buffer, no separate transpose pass, and the stores of all CTAs keep all NVLink lanes busy.
//   forward : local spectral [z][yl][x] (x fastest)  ->  T of rank d = z / nzl :  [zl][x][y], y = rank·nyl + yl   (y fastest)
//   backward: local T [zl][x][y]                      ->  spectral of rank s = y / nyl : [z = rank·nzl + zl][yl][x]  (x fastest)
// Ordering between ranks is the caller's (a barrier before and after each launch, run_fft_solve_dist).
// ---------------------------------------------------------------------------------------------------------
enum { DIST_MAX_RANKS = 16 };
template <class FT>
struct TransposePutKernel {
    static constexpr int PHASES = 2;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr size_t SMEM = sizeof(Cplx<FT>) * 32 * 33;
    int nxc, nyl, nzl, R, rank;
    int forward;
    int yperm;                          // Bounded y: Makhoul-permuted positions in T (see TransposeKernel)
    Cplx<FT>* spec[DIST_MAX_RANKS];     // every rank's spectral buffer (spec[rank]: the local one)
    Cplx<FT>* T[DIST_MAX_RANKS];        // every rank's transposed buffer
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char* smem) const {
        Cplx<FT>* tile = reinterpret_cast<Cplx<FT>*>(smem);
        const int ny = nyl * R;
        const int tx = tid & 31, ty = tid >> 5;        // 32 × 8
        const int x0 = b.x * 32;
        if (forward) {
            // Consecutive z-blocks go to consecutive destinations, starting one past this rank: the CTAs in flight at any moment write to
            // every peer at once and no two ranks start on the same destination.  (With z in natural order all ranks wrote to rank 0's
            // buffer first, then to rank 1's …: an R-to-1 incast on one NVSwitch port — 8 GPUs: 88.9 ms per step instead of 77.4 with
            // NCCL's all-to-all, profiles/r02g_bench_n8.json.)
            const int yl0 = b.y * 32, d = (b.z % R + rank + 1) % R, zl = b.z / R, z = d * nzl + zl;
            if (PHASE == 0) {                          // local read, x contiguous
                for (int r = ty; r < 32; r += 8) {
                    const int x = x0 + tx, yl = yl0 + r;
                    if (x < nxc && yl < nyl) tile[r * 33 + tx] = spec[rank][((long long)z * nyl + yl) * nxc + x];
                }
            } else {                                   // peer write, y contiguous
                Cplx<FT>* dst = T[d];
                for (int r = ty; r < 32; r += 8) {
                    const int x = x0 + r, yl = yl0 + tx;
                    if (x < nxc && yl < nyl) dst[((long long)zl * nxc + x) * ny + makhoul(rank * nyl + yl, ny, yperm)] = tile[tx * 33 + r];
                }
            }
        } else {
            // likewise the y-tiles (owners) are visited starting one past this rank
            const int ytiles = (ny + 31) / 32, own = (nyl + 31) / 32;
            const int y0 = ((b.y + (rank + 1) * own) % ytiles) * 32, zl = b.z;
            if (PHASE == 0) {                          // local read, y contiguous
                for (int r = ty; r < 32; r += 8) {
                    const int x = x0 + r, y = y0 + tx;
                    if (x < nxc && y < ny) tile[tx * 33 + r] = T[rank][((long long)zl * nxc + x) * ny + makhoul(y, ny, yperm)];
                }
            } else {                                   // peer write, x contiguous
                for (int r = ty; r < 32; r += 8) {
                    const int x = x0 + tx, y = y0 + r;
                    if (x < nxc && y < ny) {
                        const int s = y / nyl, yl = y - s * nyl;
                        spec[s][((long long)(rank * nzl + zl) * nyl + yl) * nxc + x] = tile[r * 33 + tx];
                    }
                }
            }
        }
    }
};

// Bounded x or z in a slab (y) decomposition: both are local in the first stage, so the Makhoul twiddles of their DCTs are applied on
// the local layout (complex (nxc, Ny_l, Nz), x fastest) right after the forward (z, x) FFT / right before the inverse one — the
// forward and inverse halves of PoissonMidZKernel, separated by the transposed y stage.
//   forward : X[k] = ω_k V[k] + conj(ω_k) V[N-k]                 (DCT-II, FFTW REDFT10 scaling)
//   inverse : W[k] = ½ conj(ω_k) (Φ[k] - i Φ[N-k]),  Φ[N] := 0   (DCT-III with its 1/2N folded in)
// One thread per line and reflection orbit {k, N-k}.  Line n starts at (n % inner) + (n / inner)·souter and its wavenumbers are sk
// elements apart: z lines — sk = inner = the (x, y) plane; x lines — sk = 1, inner = 1, souter = the row length (threads then run
// along k: kfast).
template <class FT>
struct TwiddleKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    long long sk, souter, count;
    int inner, N, kfast;
    int inverse;
    Cplx<FT>* spec;
    const Cd* tw;
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        const int nk = N / 2 + 1;
        const long long t = (long long)b.x * nt + tid;
        if (t >= count * nk) return;
        const long long n = kfast ? t / nk : t % count;
        const int k0 = (int)(kfast ? t % nk : t / count), k1 = (N - k0) % N;
        Cplx<FT>* line = spec + (n % inner) + (n / inner) * souter;
        Cplx<FT>* p0 = line + sk * k0;
        Cplx<FT>* p1 = line + sk * k1;
        const Cplx<FT> s0 = *p0, s1 = *p1;
        const Cd a0{(double)s0.x, (double)s0.y}, a1{(double)s1.x, (double)s1.y};
        const Cd w0 = tw[k0], w1 = tw[k1];
        Cd e0, e1;
        if (!inverse) {
            e0 = cadd(cmul(w0, a0), cmul(cconj(w0), a1));
            e1 = cadd(cmul(w1, a1), cmul(cconj(w1), a0));
        } else {
            const Cd r0 = k0 == 0 ? Cd{0.0, 0.0} : a1;
            const Cd r1 = k1 == 0 ? Cd{0.0, 0.0} : a0;
            const Cd t0{a0.x + r0.y, a0.y - r0.x}, t1{a1.x + r1.y, a1.y - r1.x};
            const Cd h0 = cmul(cconj(w0), t0), h1 = cmul(cconj(w1), t1);
            e0 = Cd{0.5 * h0.x, 0.5 * h0.y};
            e1 = Cd{0.5 * h1.x, 0.5 * h1.y};
        }
        *p0 = Cplx<FT>{(FT)e0.x, (FT)e0.y};
        *p1 = Cplx<FT>{(FT)e1.x, (FT)e1.y};
    }
};

// Pencil decompositions: the second transpose, y <-> x among the Rx ranks of a row.  T1 = [zl][xl][y] (y whole, fastest) is cut into Rx
// chunks of nyx = Ny / Rx rows; the exchange buffer B = [r][zl][yl2][xl] holds one chunk per peer r; T2 = [zl][yl2][x] has x whole and
// fastest, x = r·Nx_l + xl, stored at its Makhoul-permuted position when x is Bounded (the line the DCT's FFT runs on).
//   mode 0: B <- T1 (pack for the forward exchange)     mode 1: T2 <- B (unpack)
//   mode 2: B <- T2 (pack for the way back)             mode 3: T1 <- B (unpack)
template <class FT>
struct PencilXKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    int nxl, nyx, nzl, Rx, xperm, mode;
    Cplx<FT>* B;
    Cplx<FT>* T;
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        const long long n = (long long)b.x * nt + tid, total = (long long)nxl * nyx * nzl * Rx;
        if (n >= total) return;
        long long q = n;
        const int xl = (int)(q % nxl); q /= nxl;
        const int yl2 = (int)(q % nyx); q /= nyx;
        const int zl = (int)(q % nzl);
        const int r = (int)(q / nzl);
        const int Ny = nyx * Rx, Nx = nxl * Rx;
        Cplx<FT>* t;
        if (mode == 0 || mode == 3) t = T + ((long long)zl * nxl + xl) * Ny + r * nyx + yl2;
        else t = T + ((long long)zl * nyx + yl2) * Nx + makhoul(r * nxl + xl, Nx, xperm);
        if (mode == 0 || mode == 2) B[n] = *t; else *t = B[n];
    }
};

// Modes 0 and 3 of PencilXKernel transpose x against y: through a 32 × 33 shared-memory tile both sides are read and written in
// contiguous runs (the plain kernel strides one of them by Ny).  pack = 1: B <- T1 (mode 0); pack = 0: T1 <- B (mode 3).
template <class FT>
struct PencilYXKernel {
    static constexpr int PHASES = 2;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    static constexpr size_t SMEM = sizeof(Cplx<FT>) * 32 * 33;
    int nxl, nyx, nzl, Rx, pack;
    Cplx<FT>* B;
    Cplx<FT>* T;
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char* smem) const {
        Cplx<FT>* tile = reinterpret_cast<Cplx<FT>*>(smem);
        (void)nt;
        const int tx = tid & 31, ty = tid >> 5;        // 32 × 8
        const int x0 = b.x * 32, y0 = b.y * 32, zl = b.z % nzl, r = b.z / nzl;
        const long long Ny = (long long)nyx * Rx;
        if ((PHASE == 0) == (pack != 0)) {
            // the T1 side (y contiguous): read in phase 0 when packing, written in phase 1 when unpacking
            for (int rr = ty; rr < 32; rr += 8) {
                const int xl = x0 + rr, yl2 = y0 + tx;
                if (xl < nxl && yl2 < nyx) {
                    Cplx<FT>* p = T + ((long long)zl * nxl + xl) * Ny + (long long)r * nyx + yl2;
                    if (pack) tile[rr * 33 + tx] = *p; else *p = tile[rr * 33 + tx];
                }
            }
        } else {
            // the exchange-buffer side (x contiguous)
            for (int rr = ty; rr < 32; rr += 8) {
                const int xl = x0 + tx, yl2 = y0 + rr;
                if (xl < nxl && yl2 < nyx) {
                    Cplx<FT>* p = B + (((long long)r * nzl + zl) * nyx + yl2) * nxl + xl;
                    if (pack) *p = tile[tx * 33 + rr]; else tile[tx * 33 + rr] = *p;
                }
            }
        }
    }
};

// the divide in the pencil layout T2 = [zl][yl2][x]: global y = y0 + yl2, kz = kz0 + zl
template <class FT>
struct PoissonDividePencilKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    int Nx, nyx, nzl, y0, kz0;
    Cplx<FT>* T;
    const double* lam[3];
    double norm;
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        const int x = b.x * nt + tid, yl2 = b.y, zl = b.z;
        if (x >= Nx) return;
        const int y = y0 + yl2, kz = kz0 + zl;
        Cplx<FT>* p = T + ((long long)zl * nyx + yl2) * Nx + x;
        const Cplx<FT> e = *p;
        double s = -norm / (lam[0][x] + lam[1][y] + lam[2][kz]);
        if (x == 0 && y == 0 && kz == 0) s = 0.0;
        *p = Cplx<FT>{(FT)((double)e.x * s), (FT)((double)e.y * s)};
    }
};

// spectral divide in the transposed layout T = [zl][x][y]: global kz = rank·nzl + zl.  Bounded y (tw != nullptr): y is whole here, so
// the twiddles of its DCT wrap the divide — one thread per reflection orbit {y, Ny-y}, like PoissonMidZKernel does for z.
template <class FT>
struct PoissonDivideTKernel {
    static constexpr int PHASES = 1;
    static constexpr int THREADS = 256;
    static constexpr int MIN_BLOCKS = 1;
    int nxc, ny, nzl, kz0;
    int zl0;
    Cplx<FT>* T;
    const double* lam[3];      // λx[nxc…], λy[ny] (GLOBAL y), λz[Nz] (global z)
    const Cd* tw;              // ω_k = exp(-iπk/2Ny) of a Bounded y, else nullptr
    double norm;
    template <int PHASE>
    OC_HD void run(const Block& b, int tid, int nt, char*) const {
        const int y = b.x * nt + tid, x = b.y, zl = zl0 + b.z;
        const int kz = kz0 + zl;
        Cplx<FT>* row = T + ((long long)zl * nxc + x) * ny;
        if (!tw) {
            if (y >= ny) return;
            const double l = lam[0][x] + lam[1][y] + lam[2][kz];
            Cplx<FT> e = row[y];
            double s = -norm / l;
            if (x == 0 && y == 0 && kz == 0) s = 0.0;
            row[y] = Cplx<FT>{(FT)((double)e.x * s), (FT)((double)e.y * s)};
            return;
        }
        if (y > ny / 2) return;
        const int y1 = (ny - y) % ny;
        const Cplx<FT> s0 = row[y], s1 = row[y1];
        const Cd a0{(double)s0.x, (double)s0.y}, a1{(double)s1.x, (double)s1.y};
        const Cd w0 = tw[y], w1 = tw[y1];
        Cd X0 = cadd(cmul(w0, a0), cmul(cconj(w0), a1));
        Cd X1 = cadd(cmul(w1, a1), cmul(cconj(w1), a0));
        double c0 = -norm / (lam[0][x] + lam[1][y] + lam[2][kz]), c1 = -norm / (lam[0][x] + lam[1][y1] + lam[2][kz]);
        if (x == 0 && kz == 0) { if (y == 0) c0 = 0.0; if (y1 == 0) c1 = 0.0; }
        X0 = Cd{X0.x * c0, X0.y * c0};
        X1 = Cd{X1.x * c1, X1.y * c1};
        const Cd r0 = y == 0 ? Cd{0.0, 0.0} : X1;
        const Cd r1 = y1 == 0 ? Cd{0.0, 0.0} : X0;
        const Cd t0{X0.x + r0.y, X0.y - r0.x}, t1{X1.x + r1.y, X1.y - r1.x};
        const Cd h0 = cmul(cconj(w0), t0), h1 = cmul(cconj(w1), t1);
        row[y] = Cplx<FT>{(FT)(0.5 * h0.x), (FT)(0.5 * h0.y)};
        row[y1] = Cplx<FT>{(FT)(0.5 * h1.x), (FT)(0.5 * h1.y)};
    }
};

// ---------------------------------------------------------------------------------------------------------
// local transforms of the distributed solve: 2-D (z, x) real-to-complex batched over y, and 1-D y complex.
// ---------------------------------------------------------------------------------------------------------
template <class FT>
class DistFft {
public:
    int Nx = 0, Nyl = 0, Nz = 0, R = 1, nxc = 0, nxr = 0, Ny = 0, Nzl = 0;
    size_t work_bytes = 0;

    int C = 1;                 // the y stage runs in C sub-chunks of Nzl / C levels (pipelined against the all-to-alls)
    bool zlines = false;       // the first stage transforms z alone (pencils, Flat x)
    bool c2c = false;          // Bounded x: the (z, x) stage is complex-to-complex over whole rows (the permuted line of the DCT), like
                               // Fft3 on one GPU (plan_transforms.jl:16-136); otherwise real-to-complex, Nx/2+1 coefficients per row
    int Rx = 1;                // pencils: ranks along x — x is not local: z alone is transformed first (z()), x last (x()), all complex
    bool xonly = false;        // vertically stretched grid on slabs: z is not transformed — the first stage is the x rows alone
    bool noz = false;          // vertically stretched grid (slabs or pencils): z() does nothing
    std::string init(int nx, int nyl, int nz, int r, bool x_bounded, Stream stream, Stream ystream, int rx = 1, bool zbatch = false) {
        Nx = nx; Nyl = nyl; Nz = nz; R = r; Rx = rx; xonly = zbatch && Rx == 1; noz = zbatch;
        c2c = x_bounded || Rx > 1 || Nx == 1;      // (a Flat x: one complex number per row)
        nxc = c2c ? Nx : Nx / 2 + 1; nxr = 2 * nxc; Ny = Nyl * R; Nzl = Nz / R;
        C = Nzl % 4 == 0 ? 4 : (Nzl % 2 == 0 ? 2 : 1);
#ifndef OC_HOSTSIM
        const bool dbl = sizeof(FT) == 8;
        size_t w[3] = {0, 0, 0};
        for (cufftHandle* h : {&fwd_, &inv_, &y_}) {
            if (cufftCreate(h) != CUFFT_SUCCESS) return "cufftCreate failed";
            cufftSetAutoAllocation(*h, 0);
        }
        int n2[2] = {Nz, Nx};
        int rembed[2] = {Nz, nxr * Nyl}, cembed[2] = {Nz, nxc * Nyl};
        zlines = (Rx > 1 || Nx == 1) && !xonly;
        if (xonly) {
            // rows of Nx reals -> Nx/2+1 coefficients (or Nx complex numbers, Bounded x), one per (y, z): contiguous
            int nxv[1] = {Nx};
            const int rows = Nyl * Nz;
            if (c2c) {
                if (cufftMakePlanMany(fwd_, 1, nxv, nullptr, 1, nxc, nullptr, 1, nxc, dbl ? CUFFT_Z2Z : CUFFT_C2C, rows, &w[0]) != CUFFT_SUCCESS)
                    return "cufftMakePlanMany(x rows, complex) failed";
            } else {
                int remb[1] = {nxr}, cemb[1] = {nxc};
                if (cufftMakePlanMany(fwd_, 1, nxv, remb, 1, nxr, cemb, 1, nxc, dbl ? CUFFT_D2Z : CUFFT_R2C, rows, &w[0]) != CUFFT_SUCCESS)
                    return "cufftMakePlanMany(x rows forward) failed";
                if (cufftMakePlanMany(inv_, 1, nxv, cemb, 1, nxc, remb, 1, nxr, dbl ? CUFFT_Z2D : CUFFT_C2R, rows, &w[1]) != CUFFT_SUCCESS)
                    return "cufftMakePlanMany(x rows inverse) failed";
            }
        } else if (zlines) {
            // z lines of the local (Nx_l, Ny_l, Nz) buffer: stride = the (x, y) plane, one line per element of the plane (pencils; a Flat x
            // on slabs, where the (z, x) stage has nothing to do along x).  A Flat z (Nz = 1) needs no transform at all.
            int nzv[1] = {Nz}, emb[1] = {Nz};
            const int plane = nxc * Nyl;
            if (Nz > 1 && !noz && cufftMakePlanMany(fwd_, 1, nzv, emb, plane, 1, emb, plane, 1, dbl ? CUFFT_Z2Z : CUFFT_C2C, plane, &w[0]) != CUFFT_SUCCESS)
                return "cufftMakePlanMany(z lines) failed";
        }
        if (xonly) {
            // (nothing else: the y plan follows)
        } else if (Rx > 1) {
            // x lines of T2 = [zl][yl2][x]: contiguous, Nx·Rx long
            int nxv[1] = {Nx * Rx};
            if (cufftMakePlanMany(inv_, 1, nxv, nullptr, 1, Nx * Rx, nullptr, 1, Nx * Rx, dbl ? CUFFT_Z2Z : CUFFT_C2C, (Nz / R) * (Nyl * R / Rx), &w[1]) != CUFFT_SUCCESS)
                return "cufftMakePlanMany(x lines) failed";
        } else if (zlines) {
            // (Flat x on slabs: the z lines above are the whole first stage)
        } else if (c2c) {
            if (cufftMakePlanMany(fwd_, 2, n2, cembed, 1, nxc, cembed, 1, nxc, dbl ? CUFFT_Z2Z : CUFFT_C2C, Nyl, &w[0]) != CUFFT_SUCCESS)
                return "cufftMakePlanMany(zx complex) failed";
        } else {
        if (cufftMakePlanMany(fwd_, 2, n2, rembed, 1, nxr, cembed, 1, nxc, dbl ? CUFFT_D2Z : CUFFT_R2C, Nyl, &w[0]) != CUFFT_SUCCESS)
            return "cufftMakePlanMany(zx forward) failed";
        if (cufftMakePlanMany(inv_, 2, n2, cembed, 1, nxc, rembed, 1, nxr, dbl ? CUFFT_Z2D : CUFFT_C2R, Nyl, &w[1]) != CUFFT_SUCCESS)
            return "cufftMakePlanMany(zx inverse) failed";
        }
        int n1[1] = {Ny};
        if (cufftMakePlanMany(y_, 1, n1, nullptr, 1, Ny, nullptr, 1, Ny, dbl ? CUFFT_Z2Z : CUFFT_C2C, nxc * (Nzl / C), &w[2]) != CUFFT_SUCCESS)
            return "cufftMakePlanMany(y) failed";
        work_bytes = std::max(w[0], std::max(w[1], w[2]));
        if (work_bytes) {
            if (cudaMalloc(&work_, work_bytes) != cudaSuccess) return "cudaMalloc(cuFFT work area) failed";
            for (cufftHandle h : {fwd_, inv_, y_}) cufftSetWorkArea(h, work_);
        }
        cufftSetStream(fwd_, stream); cufftSetStream(inv_, stream); cufftSetStream(y_, ystream);
        planned_ = true;
#else
        (void)stream; (void)ystream;
#endif
        return "";
    }
    ~DistFft() {
#ifndef OC_HOSTSIM
        if (planned_) { cufftDestroy(fwd_); cufftDestroy(inv_); cufftDestroy(y_); }
        if (work_) cudaFree(work_);
#endif
    }

#ifndef OC_HOSTSIM
    void set_y_stream(Stream s) { cufftSetStream(y_, s); }
    // pencils: the z lines of the local buffer (plan fwd_) and the x lines of T2 (plan inv_), complex, in place
    std::string lines(cufftHandle h, void* buf, bool fwd) {
        cufftResult r;
        if (sizeof(FT) == 8) r = cufftExecZ2Z(h, (cufftDoubleComplex*)buf, (cufftDoubleComplex*)buf, fwd ? CUFFT_FORWARD : CUFFT_INVERSE);
        else r = cufftExecC2C(h, (cufftComplex*)buf, (cufftComplex*)buf, fwd ? CUFFT_FORWARD : CUFFT_INVERSE);
        return r == CUFFT_SUCCESS ? "" : "cuFFT line transform failed with code " + std::to_string((int)r);
    }
    std::string z(void* buf, bool fwd) { return (Nz > 1 && !noz) ? lines(fwd_, buf, fwd) : std::string(); }
    std::string x(void* buf, bool fwd) { return lines(inv_, buf, fwd); }
    std::string zx(void* buf, bool fwd) {
        if (zlines) return z(buf, fwd);
        cufftResult r;
        if (c2c) {
            if (sizeof(FT) == 8) r = cufftExecZ2Z(fwd_, (cufftDoubleComplex*)buf, (cufftDoubleComplex*)buf, fwd ? CUFFT_FORWARD : CUFFT_INVERSE);
            else r = cufftExecC2C(fwd_, (cufftComplex*)buf, (cufftComplex*)buf, fwd ? CUFFT_FORWARD : CUFFT_INVERSE);
        }
        else if (sizeof(FT) == 8) r = fwd ? cufftExecD2Z(fwd_, (cufftDoubleReal*)buf, (cufftDoubleComplex*)buf) : cufftExecZ2D(inv_, (cufftDoubleComplex*)buf, (cufftDoubleReal*)buf);
        else r = fwd ? cufftExecR2C(fwd_, (cufftReal*)buf, (cufftComplex*)buf) : cufftExecC2R(inv_, (cufftComplex*)buf, (cufftReal*)buf);
        return r == CUFFT_SUCCESS ? "" : "cuFFT zx exec failed with code " + std::to_string((int)r);
    }
    // the y transforms of sub-chunk c (levels c·Nzl/C … of T)
    std::string y(void* Tbase, bool fwd, int c) {
        char* T = (char*)Tbase + (size_t)c * (Nzl / C) * nxc * Ny * 2 * sizeof(FT);
        cufftResult r;
        if (sizeof(FT) == 8) r = cufftExecZ2Z(y_, (cufftDoubleComplex*)T, (cufftDoubleComplex*)T, fwd ? CUFFT_FORWARD : CUFFT_INVERSE);
        else r = cufftExecC2C(y_, (cufftComplex*)T, (cufftComplex*)T, fwd ? CUFFT_FORWARD : CUFFT_INVERSE);
        return r == CUFFT_SUCCESS ? "" : "cuFFT y exec failed with code " + std::to_string((int)r);
    }
#else
    // naive DFTs (test-only)
    static void dft(std::vector<Cd>& line, bool fwd) {
        const int n = (int)line.size();
        std::vector<Cd> out(n);
        const double sgn = fwd ? -1.0 : 1.0;
        for (int q = 0; q < n; ++q) {
            Cd s{0, 0};
            for (int m = 0; m < n; ++m) {
                double ang = sgn * 2.0 * M_PI * (double)((long long)q * m % n) / n;
                s = cadd(s, cmul(line[m], Cd{std::cos(ang), std::sin(ang)}));
            }
            out[q] = s;
        }
        line = out;
    }
    std::string zx(void* bufv, bool fwd) {
        FT* buf = (FT*)bufv;
        for (int j = 0; j < Nyl; ++j) {
            std::vector<Cd> full((size_t)Nx * Nz);
            auto at = [&](int i, int k) -> Cd& { return full[(size_t)i + (size_t)Nx * k]; };
            if (c2c) {
                for (int k = 0; k < Nz; ++k) for (int i = 0; i < Nx; ++i) { long long c = i + (long long)nxc * (j + (long long)Nyl * k); at(i, k) = Cd{(double)buf[2 * c], (double)buf[2 * c + 1]}; }
            } else if (fwd) {
                for (int k = 0; k < Nz; ++k) for (int i = 0; i < Nx; ++i) at(i, k) = Cd{(double)buf[i + (long long)nxr * (j + (long long)Nyl * k)], 0.0};
            } else {
                for (int k = 0; k < Nz; ++k) for (int i = 0; i < Nx; ++i) {
                    if (i < nxc) { long long c = i + (long long)nxc * (j + (long long)Nyl * k); at(i, k) = Cd{(double)buf[2 * c], (double)buf[2 * c + 1]}; }
                    else { long long c = (Nx - i) + (long long)nxc * (j + (long long)Nyl * (xonly ? k : (Nz - k) % Nz)); at(i, k) = Cd{(double)buf[2 * c], -(double)buf[2 * c + 1]}; }
                }
            }
            for (int k = 0; k < Nz; ++k) { std::vector<Cd> l(Nx); for (int i = 0; i < Nx; ++i) l[i] = at(i, k); dft(l, fwd); for (int i = 0; i < Nx; ++i) at(i, k) = l[i]; }
            if (!xonly) for (int i = 0; i < Nx; ++i) { std::vector<Cd> l(Nz); for (int k = 0; k < Nz; ++k) l[k] = at(i, k); dft(l, fwd); for (int k = 0; k < Nz; ++k) at(i, k) = l[k]; }
            if (fwd || c2c) {
                for (int k = 0; k < Nz; ++k) for (int i = 0; i < nxc; ++i) { long long c = i + (long long)nxc * (j + (long long)Nyl * k); buf[2 * c] = (FT)at(i, k).x; buf[2 * c + 1] = (FT)at(i, k).y; }
            } else {
                for (int k = 0; k < Nz; ++k) for (int i = 0; i < Nx; ++i) buf[i + (long long)nxr * (j + (long long)Nyl * k)] = (FT)at(i, k).x;
            }
        }
        return "";
    }
    // pencils (test-only naive DFTs): z lines of the local buffer, x lines of T2
    std::string z(void* bufv, bool fwd) {
        if (noz) return "";
        FT* buf = (FT*)bufv;
        const long long plane = (long long)nxc * Nyl;
        for (long long n = 0; n < plane; ++n) {
            std::vector<Cd> l(Nz);
            for (int k = 0; k < Nz; ++k) l[k] = Cd{(double)buf[2 * (n + plane * k)], (double)buf[2 * (n + plane * k) + 1]};
            dft(l, fwd);
            for (int k = 0; k < Nz; ++k) { buf[2 * (n + plane * k)] = (FT)l[k].x; buf[2 * (n + plane * k) + 1] = (FT)l[k].y; }
        }
        return "";
    }
    std::string x(void* Tv, bool fwd) {
        FT* T = (FT*)Tv;
        const int n = Nx * Rx;
        for (long long b = 0; b < (long long)Nzl * (Ny / Rx); ++b) {
            std::vector<Cd> l(n);
            for (int i = 0; i < n; ++i) l[i] = Cd{(double)T[2 * (b * n + i)], (double)T[2 * (b * n + i) + 1]};
            dft(l, fwd);
            for (int i = 0; i < n; ++i) { T[2 * (b * n + i)] = (FT)l[i].x; T[2 * (b * n + i) + 1] = (FT)l[i].y; }
        }
        return "";
    }
    std::string y(void* Tv, bool fwd, int c) {
        FT* T = (FT*)Tv + (size_t)c * (Nzl / C) * nxc * Ny * 2;
        for (long long b = 0; b < (long long)nxc * (Nzl / C); ++b) {
            std::vector<Cd> l(Ny);
            for (int y = 0; y < Ny; ++y) l[y] = Cd{(double)T[2 * (b * Ny + y)], (double)T[2 * (b * Ny + y) + 1]};
            dft(l, fwd);
            for (int y = 0; y < Ny; ++y) { T[2 * (b * Ny + y)] = (FT)l[y].x; T[2 * (b * Ny + y) + 1] = (FT)l[y].y; }
        }
        return "";
    }
#endif

private:
#ifndef OC_HOSTSIM
    cufftHandle fwd_ = 0, inv_ = 0, y_ = 0;
    void* work_ = nullptr;
    bool planned_ = false;
#endif
};

}  // namespace oc
