// oc_model.h — the model object behind the C ABI: device memory, launch sequencing, time stepping.
//
// Host-side orchestration that replaces time_step! (src/TimeSteppers/runge_kutta_3.jl:93-170,
// quasi_adams_bashforth_2.jl:74-120), update_state! (src/Models/NonhydrostaticModels/
// update_nonhydrostatic_model_state.jl:20-69), compute_pressure_correction!/make_pressure_correction!
// (pressure_correction.jl:8-53) and set!'s projection (set_nonhydrostatic_model.jl:44-57).
#pragma once
#include <cmath>
#include <limits>
#include <map>
#include <memory>
#include <tuple>
#include <stdexcept>

#include "../../include/oceananigans_b200.h"
#ifndef OC_HOSTSIM
#include <nvtx3/nvToolsExt.h>      // header-only NVTX 3: ranges cost nothing without an attached tool
#endif
#include "oc_aux.h"
#include "oc_dist.h"
#include "oc_fft.h"
#include "oc_halo.h"
#include "oc_march.h"
#include "oc_uvw.h"
#include "oc_tendency.h"

namespace oc {

// ---------------------------------------------------------------------------------------------------------
// device memory shims (product: CUDA runtime; OC_HOSTSIM: host heap, tests only)
// ---------------------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#ifndef OC_HOSTSIM
inline void cuda_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw Error(OC_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
inline void* dev_alloc(size_t bytes) {
    void* p = nullptr;
    cuda_check(cudaMalloc(&p, bytes ? bytes : 1), "cudaMalloc");
    cuda_check(cudaMemset(p, 0, bytes), "cudaMemset");
    // the memset runs on the legacy default stream, which is NOT ordered with the model's non-blocking stream
    cuda_check(cudaStreamSynchronize(cudaStreamLegacy), "cudaStreamSynchronize(legacy)");
    return p;
}
inline void dev_free(void* p) { if (p) cudaFree(p); }
inline void dev_upload(void* d, const void* h, size_t bytes, Stream s) {
    cuda_check(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s), "cudaMemcpyAsync H2D");
}
inline void dev_download(void* h, const void* d, size_t bytes, Stream s) {
    cuda_check(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H");
    cuda_check(cudaStreamSynchronize(s), "cudaStreamSynchronize");
}
// box copy between a dense host array (x fastest, extents ext) and a strided device field
inline void dev_copy_box(void* dev_origin, size_t elem, long long sy, long long sz, void* host, const int ext[3], bool to_device, Stream s) {
    if (ext[0] <= 0 || ext[1] <= 0 || ext[2] <= 0) return;
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof(p));
    // device side: pitch = sy elements, slice = sz elements  =>  rows per slice = sz / sy (exact by construction)
    cudaPitchedPtr dptr = make_cudaPitchedPtr(dev_origin, (size_t)sy * elem, (size_t)sy * elem, (size_t)(sz / sy));
    cudaPitchedPtr hptr = make_cudaPitchedPtr(host, (size_t)ext[0] * elem, (size_t)ext[0] * elem, (size_t)ext[1]);
    p.extent = make_cudaExtent((size_t)ext[0] * elem, (size_t)ext[1], (size_t)ext[2]);
    if (to_device) { p.srcPtr = hptr; p.dstPtr = dptr; p.kind = cudaMemcpyHostToDevice; }
    else { p.srcPtr = dptr; p.dstPtr = hptr; p.kind = cudaMemcpyDeviceToHost; }
    cuda_check(cudaMemcpy3DAsync(&p, s), "cudaMemcpy3DAsync");
    cuda_check(cudaStreamSynchronize(s), "cudaStreamSynchronize");
}
#else
inline void* dev_alloc(size_t bytes) { void* p = calloc(bytes ? bytes : 1, 1); if (!p) throw Error(OC_ERR_CUDA, "calloc"); return p; }
inline void dev_free(void* p) { free(p); }
inline void dev_upload(void* d, const void* h, size_t bytes, Stream) { memcpy(d, h, bytes); }
inline void dev_download(void* h, const void* d, size_t bytes, Stream) { memcpy(h, d, bytes); }
inline void dev_copy_box(void* dev_origin, size_t elem, long long sy, long long sz, void* host, const int ext[3], bool to_device, Stream) {
    for (int k = 0; k < ext[2]; ++k)
        for (int j = 0; j < ext[1]; ++j) {
            char* d = (char*)dev_origin + (size_t)(j * sy + k * sz) * elem;
            char* h = (char*)host + ((size_t)j + (size_t)ext[1] * k) * ext[0] * elem;
            if (to_device) memcpy(d, h, (size_t)ext[0] * elem); else memcpy(h, d, (size_t)ext[0] * elem);
        }
}
#endif

// One NVTX range per phase of the step (time_step! / stage / tendencies / halo / pressure solve / projection / output): Nsight
// Systems / Compute timelines show the reference's phase names (the reference has no tracing of its own beyond @info logs).
struct NvtxRange {
#ifndef OC_HOSTSIM
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
#else
    explicit NvtxRange(const char*) {}
#endif
};

struct ModelBase {
    virtual ~ModelBase() {}
    virtual void sync() = 0;
    virtual void field_info(int field, oc_field_info* info) = 0;
    virtual void transfer(int field, void* host, size_t nbytes, bool parent, bool upload) = 0;
    virtual void fill_halo_regions(const int* fields, int n, int fill_open) = 0;
    virtual void update_state(int compute_tendencies) = 0;
    virtual void compute_tendencies() = 0;
    virtual void compute_flux_bc_tendencies() = 0;
    virtual void rk3_substep(double dt, int stage) = 0;
    virtual void ab2_step(double dt, double chi) = 0;
    virtual void cache_previous_tendencies() = 0;
    virtual void compute_pressure_correction(double dt) = 0;
    virtual void make_pressure_correction(double dt) = 0;
    virtual void poisson_solve(const void* rhs, void* phi, size_t nbytes) = 0;
    virtual void set_finalize(int enforce) = 0;
    virtual void time_step_rk3(double dt) = 0;
    virtual void time_step_ab2(double dt, int euler) = 0;
    virtual void diagnostics(oc_diagnostics* out) = 0;
    virtual double field_maximum_abs(int field) = 0;
    virtual int output_begin(int field, const int lo[3], const int n[3], void* host, size_t nbytes) = 0;
    virtual int upload_begin(int field, const void* host, size_t nbytes) = 0;
    virtual void output_wait(int ticket) = 0;
    virtual bool output_test(int ticket) = 0;
    virtual void set_bc_array(int field, int side, const void* host, size_t nbytes) = 0;
    virtual void restore_previous_tendency(int field, const void* host, size_t nbytes) = 0;
    virtual void set_diffusivity_bc(int field, int side, int kind, double value) = 0;
    virtual void dist_attach(Transport* t) = 0;
    virtual int dist_rank() const = 0;
    virtual int dist_nranks() const = 0;
    virtual void timers_enable(int on) = 0;
    virtual void timers_reset() = 0;
    virtual void timers_get(double* ms, int64_t* n) = 0;
    virtual void stopwatch_start() = 0;
    virtual double stopwatch_stop() = 0;
    virtual void recover() = 0;             // after an exception crossed an entry point: back to the main stream, side streams joined
    oc_clock clock{0.0, 0, 1, INFINITY, INFINITY};
    int device = 0;                         // the CUDA device every entry point makes current for its duration (DeviceGuard)
    int64_t launches = 0;
    int64_t device_bytes = 0;
};

template <class FT>
class Model : public ModelBase {
public:
    explicit Model(const oc_config& c);
    ~Model() override;
    void sync() override;
    void field_info(int field, oc_field_info* info) override;
    void transfer(int field, void* host, size_t nbytes, bool parent, bool upload) override;
    void fill_halo_regions(const int* fields, int n, int fill_open) override;
    void update_state(int compute_tendencies) override;
    void compute_tendencies() override;
    void compute_flux_bc_tendencies() override;
    void rk3_substep(double dt, int stage) override;
    void ab2_step(double dt, double chi) override;
    void cache_previous_tendencies() override;
    void compute_pressure_correction(double dt) override;
    void make_pressure_correction(double dt) override;
    void poisson_solve(const void* rhs, void* phi, size_t nbytes) override;
    void set_finalize(int enforce) override;
    void time_step_rk3(double dt) override;
    void time_step_ab2(double dt, int euler) override;
    void diagnostics(oc_diagnostics* out) override;
    double field_maximum_abs(int field) override;
    int output_begin(int field, const int lo[3], const int n[3], void* host, size_t nbytes) override;
    int upload_begin(int field, const void* host, size_t nbytes) override;
    void output_wait(int ticket) override;
    bool output_test(int ticket) override;
    struct OutputSlot { FT* stage = nullptr; size_t cap = 0; void* ev_snap = nullptr; void* ev_done = nullptr; bool busy = false; };
    std::vector<OutputSlot> out_slots_;
    Stream out_stream_ = 0;       // device-to-host copies of output snapshots, concurrent with the time stepping on stream_
    Stream in_stream_ = 0;        // host-to-device copies of oc_upload_begin
    int acquire_slot(size_t nbytes);
    void set_bc_array(int field, int side, const void* host, size_t nbytes) override;
    void apply_flux_arrays(int f, FT* Gn, FT* Unew, FT coef);
    FT* bc_array_[OC_MAX_FIELDS][6] = {};      // device arrays of array-valued Flux BCs (nullptr: scalar)
    void restore_previous_tendency(int field, const void* host, size_t nbytes) override;
    void set_diffusivity_bc(int field, int side, int kind, double value) override;
    void recover() override;
    void dist_attach(Transport* t) override;
    int dist_rank() const override { return grank(rx_, rank_); }
    int dist_nranks() const override { return R_ * Rx_; }
    void timers_enable(int on) override { timing_ = on != 0; }
    void timers_reset() override;
    void timers_get(double* ms, int64_t* n) override;
    void stopwatch_start() override;
    double stopwatch_stop() override;

private:
    struct FieldRec {
        FT* base = nullptr;   // allocation
        FT* p = nullptr;      // interior origin
        int face[3] = {0, 0, 0};
        SideBC bc[6];
    };
    oc_config cfg_;
    Geom<FT> g_{};
    AdvCoef<FT> C_{};
    int F_ = 3;                       // prognostic fields
    int Hcfg_[3];
    size_t field_elems_ = 0;
    long long origin_off_ = 0;
    std::vector<FieldRec> state_, next_, Gn_, Gm_;
    FieldRec pNHS_, pHY_, nu_e_;
    std::vector<FieldRec> kappa_e_;
    bool has_pHY_ = false, has_amd_ = false, has_smag_ = false, has_eddy_ = false;   // has_eddy_: νₑ / κₑ fields exist (AMD, Smagorinsky, or array-valued ν / κ)
    bool hi_adv_ = false;         // WENO(7) / WENO(9) somewhere: AdvCoef::hi points to hi_tab_
    FT* hi_tab_ = nullptr;
    bool array_diff_ = false;     // ScalarDiffusivity with array-valued coefficients: the caller owns the contents of νₑ / κₑ
    bool tend_valid_ = false;     // Gⁿ == G(current state)
    // CUDA Graphs for launch-bound (small) grids: one whole time step — ~40 launches, each shorter than its launch overhead below ~10⁶ cells
    // — is captured once per buffer parity and replayed with ONE launch.  The host-side bookkeeping of the step (pointer swaps of the
    // double-buffered state and of Gⁿ / G⁻, validity flags, the clock) runs every time; in replay mode only the enqueues are skipped.
    struct GraphEntry { void* exec = nullptr; int seen = 0; };
    std::map<std::string, GraphEntry> graphs_;
    bool replay_ = false;
    bool graph_eligible() const;
    void graphs_clear();
    template <class Body> void run_graphed(const std::string& key, Body&& body);
    bool gn_pending_ = false;     // the Gⁿ slot holds the evaluation the last fused stage consumed: it becomes G⁻ (pointer swap) before the next evaluation
    bool aux_valid_ = false;      // pHY′, νₑ, κₑ computed from the current state
    struct HaloCache { HaloBox* boxes; int nboxes; int nblocks; };
    std::map<std::string, HaloCache> halo_cache_;
    Fft3<FT> fft_;
    FT* fftbuf_ = nullptr;
    double* lam_[3] = {nullptr, nullptr, nullptr};
    Cd* tw_[3] = {nullptr, nullptr, nullptr};
    HaloBox* boxes_dev_ = nullptr;
    unsigned long long* diag_dev_ = nullptr;
    Stream stream_ = 0;
    Stream stream2_ = 0;          // tracer tendencies run here, concurrently with the pressure solve on stream_
    Stream launch_stream_ = 0;    // the stream kernel launches and timer events currently go to
    void* ev_fork_ = nullptr;
    void* ev_join_ = nullptr;
    void* ev_xchg_ = nullptr;     // the deferred end-of-stage y-halo exchange on stream3_ has finished
    bool xchg_pending_ = false;   // … and the next tendencies() has not consumed it yet
    void join_exchange();
    void hydrostatic_pressure_rows(int jlo, int nj);
    void* ev_phy_ = nullptr;      // the pHY′ scan on stream2_ has finished (tendencies())
    bool tracers_in_flight_ = false;
    void fork_tracers();
    void join_tracers();
    FT gamma_[3], zeta_[3];
    // timers
    bool timing_ = false;
    struct TimerRec { int cls; void* e0; void* e1; };
    std::vector<TimerRec> timer_recs_;
    std::vector<void*> event_pool_;
    double timer_ms_[OC_TIMER_COUNT];
    int64_t timer_n_[OC_TIMER_COUNT];
    void* sw0_ = nullptr;
    void* sw1_ = nullptr;

    FieldRec alloc_field(const int face[3]);
    FieldRec& lookup(int field);
    void resolve_bcs(FieldRec& f, const oc_bc* user);
    void halo(const std::vector<FieldRec*>& fields, bool fill_open, bool defer_exchange = false);
    void aux();
    void hydrostatic_pressure();
    void compute_tendencies_if_stale();
    void rotate_pending_tendencies();
    void tendencies(int mode, double dt, int stage, double chi, bool euler, bool add_flux_bcs, bool swap_state, bool defer_tracer_join = false);
    template <int KIND> void launch_tendency(int fidx, TendencyArgs<FT>& a);
    void launch_uvw(const TendencyArgs<FT>& au, const TendencyArgs<FT>& av, const TendencyArgs<FT>& aw);
    bool uvw_ok_ = false;         // Centered(2) without Bounded dimensions, eddy closures or slabs: one launch for u, v and w (oc_uvw.h)
    enum { PART_ALL = 0, PART_INTERIOR = 1, PART_STRIPS = 2 };
    template <int KIND> void launch_march_tendency(int fidx, TendencyArgs<FT>& a, int part = PART_ALL);
    TileSrc<FT> tile_src(const FT* base, int bx, int by);
    // domain decomposition (oc_dist.h): slabs in y, Partition(1, R), or pencils, Partition(Rx, Ry).  rank_ / R_ are this rank's index
    // and the number of ranks ALONG Y (the whole job for slabs); rx_ / Rx_ along x; messages name their peers by global rank
    bool dist_ = false;
    int rank_ = 0, R_ = 1;
    int rx_ = 0, Rx_ = 1;
    int grank(int rx, int ry) const { return rx * R_ + ry; }      // x-major, like rank2index (distributed_architectures.jl:354-362)
    void exchange_x(const std::vector<FieldRec*>& fields);
    void all_to_all_x(FT* send, FT* recv);
    void run_fft_solve_pencil();
    void run_fft_solve_dist_tridiagonal();
    std::unique_ptr<Transport> transport_;
    DistFft<FT> dfft_;
    FT* distT_ = nullptr;          // transposed spectral buffer (y fastest)
    FT* diststage_ = nullptr;      // all-to-all staging
    bool p2p_ = false;             // the transposes go through peer memory (TransposePutKernel) instead of NCCL all-to-alls
    std::vector<void*> peer_spec_, peer_T_;     // every rank's fftbuf_ / distT_ (CUDA IPC mappings)
    void run_fft_solve_p2p();
    void local_twiddles(bool inverse);
    FT* halo_send_ = nullptr;
    FT* halo_recv_ = nullptr;
    size_t halo_buf_elems_ = 0;
    void exchange_y(const std::vector<FieldRec*>& fields);
    void all_to_all(FT* send, FT* recv, int c, int C);
    Stream stream3_ = 0;          // the y stage of the distributed solve (pipelined against the all-to-alls on stream_)
    std::vector<void*> ev_a2a_, ev_mid_;
    void run_fft_solve_dist();
    int xpad_ = 0;
    bool march_ok_ = false;       // no Flat dimension: the z-marching TMA kernel applies
    // vertically stretched grid: level tables (Geom::dzc …) and the FourierTridiagonalPoissonSolver's elimination factors
    bool stretched_ = false;
    FT* ztab_ = nullptr;          // six tables of N[2] + 2 H[2] + 3 levels
    FT* tri_R_ = nullptr;
    FT* tri_T_ = nullptr;
    void build_z_tables(const double* faces);
#ifndef OC_HOSTSIM
    std::map<std::tuple<const void*, int, int>, TileSrc<FT>> tmap_cache_;
#endif
    void pressure_solve_from_state();
    void run_fft_solve();
    void projection(double dt);
    void stage(int mode, double dt, int stage_no, double stage_dt, double chi, bool euler);
    void begin_timer(int cls);
    void end_timer();
    void collect_timers();
    template <class K> void go(const K& k, Dim3 grid, size_t smem, int cls);
    Dim3 grid_xyz(int threads) const { Dim3 d; d.x = (g_.N[0] + threads - 1) / threads; d.y = g_.N[1]; d.z = g_.N[2]; return d; }
};

}  // namespace oc
