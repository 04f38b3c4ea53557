// oc_model.cu — the extern "C" entry points of include/oceananigans_b200.h.
// (Model<FT>'s member definitions: oc_model_impl.h, instantiated in oc_model_f64.cu / oc_model_f32.cu)
#include "oc_model.h"

namespace oc {
extern template class Model<double>;
extern template class Model<float>;
}  // namespace oc


// =========================================================================================================
// C ABI
// =========================================================================================================
struct oc_model {
    std::unique_ptr<oc::ModelBase> impl;
};

static thread_local std::string g_last_error;

// Every entry point runs with the model's device current and restores the caller's (torch, another model on another GPU) on return.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
#ifndef OC_HOSTSIM
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev); else prev = -1;
#else
        (void)dev;
#endif
    }
    ~DeviceGuard() {
#ifndef OC_HOSTSIM
        if (prev >= 0) cudaSetDevice(prev);
#endif
    }
};

template <class Fn>
static int guarded(Fn&& fn, oc::ModelBase* model = nullptr) {
    try {
        fn();
        return OC_OK;
    } catch (const oc::Error& e) {
        g_last_error = e.what();
        if (model) model->recover();
        return e.code;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        if (model) model->recover();
        return OC_ERR_INVALID;
    }
}
#define OC_REQUIRE(m) if (!(m) || !(m)->impl) { g_last_error = "null model handle"; return OC_ERR_INVALID; } \
    DeviceGuard oc_device_guard_((m)->impl->device); oc::ModelBase* oc_model_ = (m)->impl.get(); (void)oc_model_;

extern "C" {

const char* oc_last_error(void) { return g_last_error.c_str(); }
int oc_abi_version(void) { return OC_ABI_VERSION; }

void oc_config_init(oc_config* c) {
    memset(c, 0, sizeof(*c));
    c->abi_version = OC_ABI_VERSION;
    c->float_type = OC_F64;
    for (int d = 0; d < 3; ++d) { c->N[d] = 1; c->H[d] = 0; c->topology[d] = OC_FLAT; c->delta[d] = 1.0; c->extent[d] = 1.0; }
    c->advection = OC_CENTERED2;
    c->timestepper = OC_RK3;
    c->ab2_chi = 0.1;
    c->gravity = 9.80665;                 // Oceananigans.defaults.gravitational_acceleration
    c->thermal_expansion = 1.67e-4;       // LinearEquationOfState defaults   linear_equation_of_state.jl:39-40
    c->haline_contraction = 7.8e-4;
    c->tracer_T = c->tracer_S = c->tracer_b = -1;
    c->amd_has_Cb = 0; c->amd_Cb = 0.0;
    c->coriolis_gamma = 0.0; c->coriolis_radius = 6371.0e3; c->origin_z = 0.0;      // Oceananigans.defaults.planet_radius
    c->tilted_gravity = 0; c->reserved2 = 0; c->gravity_unit_vector[0] = c->gravity_unit_vector[1] = 0.0; c->gravity_unit_vector[2] = -1.0;
    c->coriolis_beta = 0.0; c->origin_y = 0.0; c->coriolis_fxyz[0] = c->coriolis_fxyz[1] = c->coriolis_fxyz[2] = 0.0;
    c->smagorinsky = 0; c->smag_C = 0.16; c->smag_Cb = 1.0;          // smagorinsky.jl:77-78, lilly_coefficient.jl:47
    for (int t = 0; t < OC_MAX_TRACERS; ++t) c->smag_Pr[t] = 1.0;
    c->amd_Cnu = 1.0 / 3.0;
    for (int t = 0; t < OC_MAX_TRACERS; ++t) c->amd_Ckappa[t] = 1.0 / 3.0;
    c->z_stretched = 0;
    c->z_faces = nullptr;
    c->has_advection_dir = 0;
    c->array_diffusivity = 0; c->dist_ranks_x = 0;
    c->advection_dir[0] = c->advection_dir[1] = c->advection_dir[2] = OC_CENTERED2;
}

int oc_model_create(const oc_config* cfg, oc_model** out) {
    if (!cfg || !out) { g_last_error = "null argument"; return OC_ERR_INVALID; }
    *out = nullptr;
    return guarded([&] {
        std::unique_ptr<oc_model> m(new oc_model);
        if (cfg->float_type == OC_F64) m->impl.reset(new oc::Model<double>(*cfg));
        else if (cfg->float_type == OC_F32) m->impl.reset(new oc::Model<float>(*cfg));
        else throw oc::Error(OC_ERR_INVALID, "float_type must be OC_F64 or OC_F32");
        *out = m.release();
    });
}
int oc_model_destroy(oc_model* m) {
    if (!m) return OC_OK;
    DeviceGuard dg(m->impl ? m->impl->device : 0);
    return guarded([&] { delete m; });
}
int oc_sync(oc_model* m) { OC_REQUIRE(m); return guarded([&] { m->impl->sync(); }, oc_model_); }
int oc_field_info_get(oc_model* m, int field, oc_field_info* info) { OC_REQUIRE(m); return guarded([&] { m->impl->field_info(field, info); }, oc_model_); }
int oc_upload_interior(oc_model* m, int field, const void* host, size_t nbytes) { OC_REQUIRE(m); return guarded([&] { m->impl->transfer(field, const_cast<void*>(host), nbytes, false, true); }, oc_model_); }
int oc_download_interior(oc_model* m, int field, void* host, size_t nbytes) { OC_REQUIRE(m); return guarded([&] { m->impl->transfer(field, host, nbytes, false, false); }, oc_model_); }
int oc_upload_parent(oc_model* m, int field, const void* host, size_t nbytes) { OC_REQUIRE(m); return guarded([&] { m->impl->transfer(field, const_cast<void*>(host), nbytes, true, true); }, oc_model_); }
int oc_download_parent(oc_model* m, int field, void* host, size_t nbytes) { OC_REQUIRE(m); return guarded([&] { m->impl->transfer(field, host, nbytes, true, false); }, oc_model_); }
int oc_fill_halo_regions(oc_model* m, const int* fields, int nfields, int fill_open_bcs) { OC_REQUIRE(m); return guarded([&] { m->impl->fill_halo_regions(fields, nfields, fill_open_bcs); }, oc_model_); }
int oc_update_state(oc_model* m, int compute_tendencies) { OC_REQUIRE(m); return guarded([&] { m->impl->update_state(compute_tendencies); }, oc_model_); }
int oc_compute_tendencies(oc_model* m) { OC_REQUIRE(m); return guarded([&] { m->impl->compute_tendencies(); }, oc_model_); }
int oc_compute_flux_bc_tendencies(oc_model* m) { OC_REQUIRE(m); return guarded([&] { m->impl->compute_flux_bc_tendencies(); }, oc_model_); }
int oc_rk3_substep(oc_model* m, double dt, int stage) { OC_REQUIRE(m); return guarded([&] { m->impl->rk3_substep(dt, stage); }, oc_model_); }
int oc_ab2_step(oc_model* m, double dt, double chi) { OC_REQUIRE(m); return guarded([&] { m->impl->ab2_step(dt, chi); }, oc_model_); }
int oc_cache_previous_tendencies(oc_model* m) { OC_REQUIRE(m); return guarded([&] { m->impl->cache_previous_tendencies(); }, oc_model_); }
int oc_compute_pressure_correction(oc_model* m, double dt) { OC_REQUIRE(m); return guarded([&] { m->impl->compute_pressure_correction(dt); }, oc_model_); }
int oc_make_pressure_correction(oc_model* m, double dt) { OC_REQUIRE(m); return guarded([&] { m->impl->make_pressure_correction(dt); }, oc_model_); }
int oc_poisson_solve(oc_model* m, const void* rhs, void* phi, size_t nbytes) { OC_REQUIRE(m); return guarded([&] { m->impl->poisson_solve(rhs, phi, nbytes); }, oc_model_); }
int oc_set_finalize(oc_model* m, int enforce) { OC_REQUIRE(m); return guarded([&] { m->impl->set_finalize(enforce); }, oc_model_); }
int oc_time_step_rk3(oc_model* m, double dt) { OC_REQUIRE(m); return guarded([&] { m->impl->time_step_rk3(dt); }, oc_model_); }
int oc_time_step_ab2(oc_model* m, double dt, int euler) { OC_REQUIRE(m); return guarded([&] { m->impl->time_step_ab2(dt, euler); }, oc_model_); }
int oc_get_clock(oc_model* m, oc_clock* c) { OC_REQUIRE(m); *c = m->impl->clock; return OC_OK; }
int oc_set_clock(oc_model* m, const oc_clock* c) { OC_REQUIRE(m); m->impl->clock = *c; return OC_OK; }
int oc_restore_previous_tendency(oc_model* m, int field, const void* host, size_t nbytes) { OC_REQUIRE(m); return guarded([&] { m->impl->restore_previous_tendency(field, host, nbytes); }, oc_model_); }
int oc_compute_diagnostics(oc_model* m, oc_diagnostics* out) { OC_REQUIRE(m); return guarded([&] { m->impl->diagnostics(out); }, oc_model_); }
int oc_set_bc_array(oc_model* m, int field, int side, const void* host, size_t nbytes) {
    OC_REQUIRE(m);
    if (!host) { g_last_error = "null argument"; return OC_ERR_INVALID; }
    return guarded([&] { m->impl->set_bc_array(field, side, host, nbytes); });
}
int oc_set_diffusivity_bc(oc_model* m, int field, int side, int kind, double value) { OC_REQUIRE(m); return guarded([&] { m->impl->set_diffusivity_bc(field, side, kind, value); }, oc_model_); }
int oc_output_begin(oc_model* m, int field, const int lo[3], const int n[3], void* host, size_t nbytes, int* ticket) {
    OC_REQUIRE(m);
    if (!lo || !n || !host || !ticket) { g_last_error = "null argument"; return OC_ERR_INVALID; }
    return guarded([&] { *ticket = m->impl->output_begin(field, lo, n, host, nbytes); });
}
int oc_upload_begin(oc_model* m, int field, const void* host, size_t nbytes, int* ticket) {
    OC_REQUIRE(m);
    if (!host || !ticket) { g_last_error = "null argument"; return OC_ERR_INVALID; }
    return guarded([&] { *ticket = m->impl->upload_begin(field, host, nbytes); }, oc_model_);
}
int oc_output_wait(oc_model* m, int ticket) { OC_REQUIRE(m); return guarded([&] { m->impl->output_wait(ticket); }, oc_model_); }
int oc_output_test(oc_model* m, int ticket, int* done) {
    OC_REQUIRE(m);
    if (!done) { g_last_error = "null argument"; return OC_ERR_INVALID; }
    return guarded([&] { *done = m->impl->output_test(ticket) ? 1 : 0; });
}
int oc_field_maximum_abs(oc_model* m, int field, double* out) {
    OC_REQUIRE(m);
    if (!out) { g_last_error = "null argument"; return OC_ERR_INVALID; }
    return guarded([&] { *out = m->impl->field_maximum_abs(field); });
}
int oc_dist_unique_id(void* id128) {
    if (!id128) { g_last_error = "null argument"; return OC_ERR_INVALID; }
    return guarded([&] {
#ifndef OC_HOSTSIM
        std::string e = oc::nccl_api().load();
        if (!e.empty()) throw oc::Error(OC_ERR_CUDA, e);
        oc::NcclApi::UniqueId id;
        int rc = oc::nccl_api().GetUniqueId(&id);
        if (rc != 0) throw oc::Error(OC_ERR_CUDA, std::string("ncclGetUniqueId: ") + oc::nccl_api().GetErrorString(rc));
        memcpy(id128, id.internal, 128);
#else
        memset(id128, 0, 128);
#endif
    });
}
int oc_dist_attach_nccl(oc_model* m, const void* id128) {
    OC_REQUIRE(m);
    return guarded([&] {
#ifndef OC_HOSTSIM
        if (!id128) throw oc::Error(OC_ERR_INVALID, "null NCCL id");
        std::unique_ptr<oc::NcclTransport> t(new oc::NcclTransport);
        std::string e = t->init(m->impl->dist_rank(), m->impl->dist_nranks(), id128);
        if (!e.empty()) throw oc::Error(OC_ERR_CUDA, e);
        m->impl->dist_attach(t.release());
#else
        (void)id128;
        throw oc::Error(OC_ERR_UNSUPPORTED, "the host simulation has no NCCL transport (use oc_dist_attach_host)");
#endif
    });
}
int oc_dist_attach_host(oc_model* m, oc_exchange_fn fn, void* user) {
    OC_REQUIRE(m);
    return guarded([&] {
#ifdef OC_HOSTSIM
        if (!fn) throw oc::Error(OC_ERR_INVALID, "null exchange callback");
        const char* threads = getenv("OC_HOSTSIM_THREADS");
        m->impl->dist_attach(new oc::HostTransport(fn, user, m->impl->dist_rank(), m->impl->dist_nranks(), threads && atoi(threads) != 0));
#else
        (void)fn; (void)user;
        throw oc::Error(OC_ERR_UNSUPPORTED, "the host-callback transport is a test facility of the host simulation build; use oc_dist_attach_nccl");
#endif
    });
}
int oc_timers_enable(oc_model* m, int enable) { OC_REQUIRE(m); return guarded([&] { m->impl->timers_enable(enable); }, oc_model_); }
int oc_timers_reset(oc_model* m) { OC_REQUIRE(m); return guarded([&] { m->impl->timers_reset(); }, oc_model_); }
int oc_timers_get(oc_model* m, double* ms, int64_t* n) { OC_REQUIRE(m); return guarded([&] { m->impl->timers_get(ms, n); }, oc_model_); }
int oc_stopwatch_start(oc_model* m) { OC_REQUIRE(m); return guarded([&] { m->impl->stopwatch_start(); }, oc_model_); }
int oc_stopwatch_stop(oc_model* m, double* ms) { OC_REQUIRE(m); return guarded([&] { *ms = m->impl->stopwatch_stop(); }, oc_model_); }
int oc_host_alloc(void** ptr, size_t nbytes) {
    return guarded([&] {
#ifndef OC_HOSTSIM
        oc::cuda_check(cudaHostAlloc(ptr, nbytes, cudaHostAllocDefault), "cudaHostAlloc");
#else
        *ptr = malloc(nbytes);
#endif
    });
}
int oc_host_free(void* ptr) {
    return guarded([&] {
#ifndef OC_HOSTSIM
        oc::cuda_check(cudaFreeHost(ptr), "cudaFreeHost");
#else
        free(ptr);
#endif
    });
}
int64_t oc_launch_count(oc_model* m) { return (m && m->impl) ? m->impl->launches : -1; }
int oc_device_bytes(oc_model* m, int64_t* bytes) { OC_REQUIRE(m); *bytes = m->impl->device_bytes; return OC_OK; }

}  // extern "C"
