// Diffusion process: ancestral DDPM sampling (START_X, FIXED_SMALL) around the denoiser.
// Replaces reference diffusion/gaussian_diffusion.py:160-197 (coefficient tables, float64 on the
// host exactly like the reference, then .float()), :253-275, :277-388, :496-548, :598-736 and
// diffusion/respace.py:64-129.
//
// Every per-step scalar lives in a device table indexed by a device-side step counter, so one
// captured CUDA graph of a step can be replayed for every timestep without host patching.  The
// counter is moved by the last kernel of a step (k_step_io, denoiser.cu), which also prepares the
// next step's decoder input, so a plain step is: decoder body (GEMMs + attention) + one tail kernel.
#include "common.cuh"

int idb_denoiser_tokens(idb_handle* h, const float* x, const long long* tstep, cudaStream_t st);
int idb_denoiser_body(idb_handle* h, cudaStream_t st);
int idb_denoiser_heads(idb_handle* h, const float* gt, const unsigned char* mask, float* out, cudaStream_t st);
int idb_step_finish(idb_handle* h, const float* x0, const float* xt, const float* noise, int tape_mode, float* x_next, int emit_next,
                    cudaStream_t st);
int idb_step_tail(idb_handle* h, const float* gt, const unsigned char* mask, float* x0_out, const float* xt, const float* noise,
                  int tape_mode, float* x_next, int emit_next, cudaStream_t st);
int idb_correction_apply_dev(idb_handle* h, float* x0, const float* gt, int t, cudaStream_t st);

struct Sampler {
    float *x_a = nullptr, *x_b = nullptr, *x0 = nullptr;
    size_t numel = 0;
    cudaGraphExec_t step_graph = nullptr;     // predict + finish for a plain step
    cudaGraphExec_t predict_graph = nullptr;  // predict only (correction steps)
    const float* g_gt = nullptr; const unsigned char* g_mask = nullptr; const float* g_tape = nullptr;
    int gB = 0, gT = 0;
    int launches_per_step = 0, launches_per_predict = 0;
};

namespace {

__global__ void k_set_counter(int* counter, int v) { *counter = v; }

}  // namespace

extern "C" int idb_diffusion_init(idb_handle* h, const double* betas, const int64_t* timestep_map, int n) {
    if (!h || !betas || n <= 0) return IDB_ERR_ARG;
    Diffusion& df = h->diff;
    // float64 tables exactly as GaussianDiffusion.__init__ (gaussian_diffusion.py:160-197)
    std::vector<double> ac(n), ac_prev(n), post_var(n), logvar(n), c1(n), c2(n);
    double prod = 1.0;
    for (int i = 0; i < n; i++) {
        if (!(betas[i] > 0 && betas[i] <= 1)) return idb_fail(h, IDB_ERR_ARG, "betas must be in (0,1]");
        prod *= (1.0 - betas[i]);
        ac[i] = prod;
        ac_prev[i] = i == 0 ? 1.0 : ac[i - 1];
    }
    for (int i = 0; i < n; i++) {
        post_var[i] = betas[i] * (1.0 - ac_prev[i]) / (1.0 - ac[i]);
        c1[i] = betas[i] * std::sqrt(ac_prev[i]) / (1.0 - ac[i]);
        c2[i] = (1.0 - ac_prev[i]) * std::sqrt(1.0 - betas[i]) / (1.0 - ac[i]);
    }
    for (int i = 0; i < n; i++) logvar[i] = std::log(i == 0 ? post_var[n > 1 ? 1 : 0] : post_var[i]);
    df.host.resize(n);
    for (int i = 0; i < n; i++) {
        StepParams& p = df.host[i];
        p.t = timestep_map ? (long long)timestep_map[i] : i;
        p.c1 = (float)c1[i];
        p.c2 = (float)c2[i];
        // reference: th.exp(0.5 * float32(logvar)) evaluated in float32
        const float lv = (float)logvar[i];
        p.sigma_nz = i == 0 ? 0.0f : expf(0.5f * lv);
        p.pad = 0.f;
    }
    if (df.tbl) cudaFree(df.tbl);
    CUDA_TRY(h, cudaMalloc((void**)&df.tbl, sizeof(StepParams) * n));
    CUDA_TRY(h, cudaMemcpy(df.tbl, df.host.data(), sizeof(StepParams) * n, cudaMemcpyHostToDevice));
    df.n = n;
    if (h->sampler) {
        if (h->sampler->step_graph) cudaGraphExecDestroy(h->sampler->step_graph);
        if (h->sampler->predict_graph) cudaGraphExecDestroy(h->sampler->predict_graph);
        h->sampler->step_graph = h->sampler->predict_graph = nullptr;
    }
    return IDB_OK;
}

static int sampler_ready(idb_handle* h, size_t numel) {
    if (!h->diff.n) return idb_fail(h, IDB_ERR_STATE, "idb_diffusion_init first");
    if (!h->den.B) return idb_fail(h, IDB_ERR_STATE, "idb_denoiser_bind first");
    if (!h->sampler) h->sampler = new Sampler();
    Sampler& s = *h->sampler;
    if (s.numel != numel) {
        if (s.x_a) { cudaFree(s.x_a); cudaFree(s.x_b); cudaFree(s.x0); }
        CUDA_TRY(h, cudaMalloc((void**)&s.x_a, numel * sizeof(float)));
        CUDA_TRY(h, cudaMalloc((void**)&s.x_b, numel * sizeof(float)));
        CUDA_TRY(h, cudaMalloc((void**)&s.x0, numel * sizeof(float)));
        s.numel = numel;
        if (s.step_graph) { cudaGraphExecDestroy(s.step_graph); s.step_graph = nullptr; }
        if (s.predict_graph) { cudaGraphExecDestroy(s.predict_graph); s.predict_graph = nullptr; }
    }
    return IDB_OK;
}

static size_t sample_numel(idb_handle* h) {
    const idb_denoiser_config& c = h->den.cfg;
    return (size_t)h->den.B * (c.c_body + c.c_obj + c.c_extra) * h->den.T;
}

// predict with the device counter already set: tokens from x_t, decoder, heads
static int predict_dev(idb_handle* h, const float* x_t, const float* gt, const unsigned char* mask, float* x0, cudaStream_t st) {
    int rc;
    if ((rc = idb_denoiser_tokens(h, x_t, nullptr, st))) return rc;
    if ((rc = idb_denoiser_body(h, st))) return rc;
    return idb_denoiser_heads(h, gt, mask, x0, st);
}

static int set_step(idb_handle* h, int i, cudaStream_t st) {
    if (!h->den.step_cur) return idb_fail(h, IDB_ERR_STATE, "idb_denoiser_bind first");
    k_set_counter<<<1, 1, 0, st>>>(h->den.step_cur, i);
    LAUNCH_CHECK(h);
    return IDB_OK;
}

extern "C" int idb_p_sample_predict(idb_handle* h, int i, const float* x_t, const float* gt, const uint8_t* mask, float* x0_out, void* stream) {
    if (!h || !x_t || !x0_out) return IDB_ERR_ARG;
    if (i < 0 || i >= h->diff.n) return idb_fail(h, IDB_ERR_ARG, "step index out of range");
    if ((gt == nullptr) != (mask == nullptr)) return idb_fail(h, IDB_ERR_ARG, "gt and mask must be given together");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = set_step(h, i, st);
    if (rc) return rc;
    return predict_dev(h, x_t, gt, mask, x0_out, st);
}

extern "C" int idb_p_sample_finish(idb_handle* h, int i, const float* x0, const float* x_t, const float* noise, float* x_out, void* stream) {
    if (!h || !x0 || !x_t || !noise || !x_out) return IDB_ERR_ARG;
    if (i < 0 || i >= h->diff.n) return idb_fail(h, IDB_ERR_ARG, "step index out of range");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = set_step(h, i, st);
    if (rc) return rc;
    return idb_step_finish(h, x0, x_t, noise, 0, x_out, 0, st);
}

extern "C" int idb_p_sample(idb_handle* h, int i, const float* x_t, const float* noise, const float* gt, const uint8_t* mask,
                            float* x_out, float* x0_out, void* stream) {
    if (!h || !x_t || !noise || !x_out) return IDB_ERR_ARG;
    if (i < 0 || i >= h->diff.n) return idb_fail(h, IDB_ERR_ARG, "step index out of range");
    if ((gt == nullptr) != (mask == nullptr)) return idb_fail(h, IDB_ERR_ARG, "gt and mask must be given together");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = set_step(h, i, st);
    if (rc) return rc;
    if ((rc = idb_denoiser_tokens(h, x_t, nullptr, st))) return rc;
    if ((rc = idb_denoiser_body(h, st))) return rc;
    return idb_step_tail(h, gt, mask, x0_out, x_t, noise, 0, x_out, 0, st);
}

extern "C" int idb_p_sample_loop(idb_handle* h, const float* tape, const float* gt, const uint8_t* mask, int correction,
                                 int use_graph, float* x_out, void* stream) {
    if (!h || !tape || !x_out) return IDB_ERR_ARG;
    if ((gt == nullptr) != (mask == nullptr)) return idb_fail(h, IDB_ERR_ARG, "gt and mask must be given together");
    const size_t numel = sample_numel(h);
    int rc = sampler_ready(h, numel);
    if (rc) return rc;
    Sampler& s = *h->sampler;
    Diffusion& df = h->diff;
    cudaStream_t st = (cudaStream_t)stream;
    const int n = df.n;
    // The per-step graphs read x from s.x_a and write s.x_b, then the roles swap: capture two
    // parities instead (a -> b, b -> a) by capturing ONE graph that does a full a -> b -> a pair?
    // Simpler and just as cheap: the graph always reads x_a and writes x_a (x0 is a separate
    // buffer, and the posterior is elementwise, so in-place is safe).
    CUDA_TRY(h, cudaMemcpyAsync(s.x_a, tape, numel * sizeof(float), cudaMemcpyDefault, st));
    if ((rc = set_step(h, n - 1, st))) return rc;
    // decoder input of the first step; every later step gets its tokens from the previous step's tail
    if ((rc = idb_denoiser_tokens(h, s.x_a, nullptr, st))) return rc;

    const bool graph_ok = use_graph != 0;
    if (graph_ok && (!s.step_graph || s.g_gt != gt || s.g_mask != mask || s.g_tape != tape || s.gB != h->den.B || s.gT != h->den.T)) {
        if (s.step_graph) { cudaGraphExecDestroy(s.step_graph); s.step_graph = nullptr; }
        if (s.predict_graph) { cudaGraphExecDestroy(s.predict_graph); s.predict_graph = nullptr; }
        cudaStream_t cs;
        CUDA_TRY(h, cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaGraph_t g;
        // full plain step
        long long saved = h->launches;
        CUDA_TRY(h, cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        rc = idb_denoiser_body(h, cs);
        if (!rc) rc = idb_step_tail(h, gt, mask, nullptr, s.x_a, tape, 1, s.x_a, 1, cs);
        cudaError_t ce = cudaStreamEndCapture(cs, &g);
        if (rc || ce != cudaSuccess) { cudaStreamDestroy(cs); return rc ? rc : idb_fail(h, IDB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce)); }
        s.launches_per_step = (int)(h->launches - saved);
        CUDA_TRY(h, cudaGraphInstantiate(&s.step_graph, g, 0));
        cudaGraphDestroy(g);
        // predict only
        CUDA_TRY(h, cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        rc = idb_denoiser_body(h, cs);
        if (!rc) rc = idb_denoiser_heads(h, gt, mask, s.x0, cs);
        ce = cudaStreamEndCapture(cs, &g);
        if (rc || ce != cudaSuccess) { cudaStreamDestroy(cs); return rc ? rc : idb_fail(h, IDB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce)); }
        s.launches_per_predict = s.launches_per_step;
        CUDA_TRY(h, cudaGraphInstantiate(&s.predict_graph, g, 0));
        cudaGraphDestroy(g);
        cudaStreamDestroy(cs);
        h->launches = saved;
        s.g_gt = gt; s.g_mask = mask; s.g_tape = tape; s.gB = h->den.B; s.gT = h->den.T;
    }

    for (int i = n - 1; i >= 0; i--) {
        const long long t = df.host[i].t;  // denoised_fn receives the UN-mapped t (gaussian_diffusion.py:356): i
        const bool corr = correction && i <= 500 && (i % 50 == 0);
        (void)t;
        if (!corr) {
            if (graph_ok) {
                CUDA_TRY(h, cudaGraphLaunch(s.step_graph, st));
                h->launches += s.launches_per_step;
            } else {
                if ((rc = idb_denoiser_body(h, st))) return rc;
                if ((rc = idb_step_tail(h, gt, mask, nullptr, s.x_a, tape, 1, s.x_a, 1, st))) return rc;
            }
        } else {
            if (graph_ok) {
                CUDA_TRY(h, cudaGraphLaunch(s.predict_graph, st));
                h->launches += s.launches_per_predict;
            } else {
                if ((rc = idb_denoiser_body(h, st))) return rc;
                if ((rc = idb_denoiser_heads(h, gt, mask, s.x0, st))) return rc;
            }
            if ((rc = idb_correction_apply_dev(h, s.x0, gt, i, st))) return rc;
            if ((rc = idb_step_finish(h, s.x0, s.x_a, tape, 1, s.x_a, 1, st))) return rc;
        }
    }
    CUDA_TRY(h, cudaMemcpyAsync(x_out, s.x_a, numel * sizeof(float), cudaMemcpyDefault, st));
    return IDB_OK;
}

void idb_sampler_drop_graphs(idb_handle* h) {
    if (!h->sampler) return;
    Sampler& s = *h->sampler;
    if (s.step_graph) { cudaGraphExecDestroy(s.step_graph); s.step_graph = nullptr; }
    if (s.predict_graph) { cudaGraphExecDestroy(s.predict_graph); s.predict_graph = nullptr; }
}

void idb_sampler_release(idb_handle* h) {
    if (h->sampler) {
        Sampler& s = *h->sampler;
        if (s.step_graph) cudaGraphExecDestroy(s.step_graph);
        if (s.predict_graph) cudaGraphExecDestroy(s.predict_graph);
        if (s.x_a) { cudaFree(s.x_a); cudaFree(s.x_b); cudaFree(s.x0); }
        delete h->sampler;
        h->sampler = nullptr;
    }
    if (h->diff.tbl) cudaFree(h->diff.tbl);
    h->diff = Diffusion();
}
