// multi.cu — ONE call from ONE process drives several GPUs (`devices[], ndev`).
//
// The reference's callers are single calls from one process (max_localize(model),
// src/localization/max_localize.jl:44-101; disentangle(model), disentangle.jl:631-728; omega(model),
// src/spread.jl:370-381).  A wb200_multi owns one k-slab context per entry of `devices` (the same
// device may appear more than once: two slabs on one GPU is how the sharded code path is tested on a
// one-GPU box), a full-size gauge array per slab that the other slabs' pack kernels read directly
// (cudaDeviceEnablePeerAccess: plain loads over NVLink, no IPC handles), and one host thread per slab
// for the duration of a call.  The slab contexts run exactly the code of the one-process-per-GPU mode
// (api.cu / optimizer.cu); their collective is an in-process all-reduce: every thread copies its
// scalars to pinned host memory, meets the others at a barrier and sums them in rank order — the
// result is bitwise identical on every slab and independent of thread timing.  A failing slab aborts
// the barrier, so its peers return an error instead of waiting forever.
#include "common.cuh"

#include <condition_variable>
#include <mutex>
#include <thread>

namespace {

struct HostBarrier {
    std::mutex m;
    std::condition_variable cv;
    int n = 1, count = 0;
    unsigned gen = 0;
    bool aborted = false;
    bool wait() {
        std::unique_lock<std::mutex> lk(m);
        if (aborted) return false;
        const unsigned g = gen;
        if (++count == n) {
            count = 0;
            gen++;
            cv.notify_all();
            return true;
        }
        cv.wait(lk, [&] { return gen != g || aborted; });
        return gen != g;
    }
    void abort() {
        std::lock_guard<std::mutex> lk(m);
        aborted = true;
        cv.notify_all();
    }
    void reset() {
        std::lock_guard<std::mutex> lk(m);
        aborted = false;
        count = 0;
    }
};

}  // namespace

struct wb200_multi {
    int n = 0, nb = 0, nw = 0, nk = 0, nn = 0;
    std::vector<int> devices, bounds;
    std::vector<wb200_ctx*> ctx;
    std::vector<cplx*> Ufull;            // per slab: [nk][nw][nb] on its device; the slab owns its rows
    std::vector<void*> flags;            // per slab: peer comm buffer
    WbInproc inproc;                     // events + host rendezvous of the slab threads
    HostBarrier bar;
    int maxc = 0;
    double* h_x[2] = {nullptr, nullptr};   // pinned [parity][slab][maxc]: contributions
    double* h_r[2] = {nullptr, nullptr};   // pinned [parity][slab][maxc]: per-slab copy of the result
    std::vector<unsigned> calls;
    std::string err;
    struct User { wb200_multi* m; int rank; };
    std::vector<User> users;
};

namespace {

thread_local std::string g_multi_create_error;

int multi_allreduce(void* dptr, int count, int op, void* stream, void* user) {
    auto* u = (wb200_multi::User*)user;
    wb200_multi* m = u->m;
    const int r = u->rank;
    if (count > m->maxc) { m->bar.abort(); return 3; }
    const unsigned p = m->calls[r]++ & 1u;
    cudaStream_t s = (cudaStream_t)stream;
    double* mine = m->h_x[p] + (size_t)r * m->maxc;
    if (cudaMemcpyAsync(mine, dptr, sizeof(double) * count, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) {
        m->bar.abort();
        return 1;
    }
    if (!m->bar.wait()) return 2;
    double* out = m->h_r[p] + (size_t)r * m->maxc;
    for (int i = 0; i < count; i++) {
        double acc = m->h_x[p][i];
        for (int q = 1; q < m->n; q++) {
            const double v = m->h_x[p][(size_t)q * m->maxc + i];
            acc = op == 0 ? acc + v : (op == 1 ? fmax(acc, v) : fmin(acc, v));
        }
        out[i] = acc;
    }
    if (cudaMemcpyAsync(dptr, out, sizeof(double) * count, cudaMemcpyHostToDevice, s) != cudaSuccess) {
        m->bar.abort();
        return 1;
    }
    return 0;
}

// f(rank) on one host thread per slab; returns the first primary failure (a slab that only saw its peers
// abort reports WB200_ERR_CUDA "all-reduce callback failed", which is secondary)
template <class F>
int run_all(wb200_multi* m, F f) {
    std::vector<int> st(m->n, WB200_OK);
    m->bar.reset();
    std::vector<std::thread> th;
    for (int r = 0; r < m->n; r++)
        th.emplace_back([&, r] {
            cudaSetDevice(m->devices[r]);
            st[r] = f(r);
            if (st[r] != WB200_OK) m->bar.abort();
        });
    for (auto& t : th) t.join();
    int first = WB200_OK;
    for (int r = 0; r < m->n; r++) {
        if (st[r] == WB200_OK) continue;
        const bool secondary = m->ctx[r]->err.find("all-reduce callback failed") != std::string::npos;
        if (first == WB200_OK || !secondary) {
            if (first == WB200_OK || !secondary) {
                first = st[r];
                m->err = "slab " + std::to_string(r) + " (device " + std::to_string(m->devices[r]) + "): " + m->ctx[r]->err;
            }
            if (!secondary) break;
        }
    }
    return first;
}

}  // namespace

extern "C" const char* wb200_multi_last_error(const wb200_multi* m) {
    return m ? m->err.c_str() : g_multi_create_error.c_str();
}

extern "C" void wb200_multi_destroy(wb200_multi* m) {
    if (!m) return;
    for (int r = 0; r < (int)m->ctx.size(); r++) {
        if (m->ctx[r]) wb200_destroy(m->ctx[r]);
        if (r < (int)m->Ufull.size() && m->Ufull[r]) {
            cudaSetDevice(m->devices[r]);
            wb_dev_free(m->Ufull[r]);
        }
        if (r < (int)m->flags.size() && m->flags[r]) {
            cudaSetDevice(m->devices[r]);
            wb_dev_free(m->flags[r]);
        }
        for (auto* v : {&m->inproc.ev_ready, &m->inproc.ev_packed, &m->inproc.ev_ar[0], &m->inproc.ev_ar[1]})
            if (r < (int)v->size() && (*v)[r]) { cudaSetDevice(m->devices[r]); cudaEventDestroy((*v)[r]); }
    }
    for (int p = 0; p < 2; p++) {
        if (m->h_x[p]) wb_pinned_free(m->h_x[p]);
        if (m->h_r[p]) wb_pinned_free(m->h_r[p]);
    }
    delete m;
}

extern "C" int wb200_multi_create(wb200_multi** out, const int* devices, int ndev, int nb, int nw, int nk, int nn,
                                  const int32_t* kpb_k, const double* bvec_cart, const double* bweights,
                                  const double* M, unsigned flags) {
    if (!out) return WB200_ERR_ARG;
    *out = nullptr;
    if (!devices || ndev <= 0 || ndev > nk || !M || !kpb_k || !bvec_cart || !bweights) {
        g_multi_create_error = "wb200_multi_create: need 1 <= ndev <= nk and non-NULL arrays";
        return WB200_ERR_ARG;
    }
    wb200_multi* m = new wb200_multi();
    m->n = ndev; m->nb = nb; m->nw = nw; m->nk = nk; m->nn = nn;
    m->devices.assign(devices, devices + ndev);
    for (int r = 0; r <= ndev; r++) m->bounds.push_back((int)((long long)nk * r / ndev));
    m->ctx.assign(ndev, nullptr);
    m->Ufull.assign(ndev, nullptr);
    m->calls.assign(ndev, 0);
    m->users.resize(ndev);
    m->bar.n = ndev;
    m->maxc = 4 * nw + 64;
    auto fail = [&](int code, const std::string& msg) {
        g_multi_create_error = msg;
        wb200_multi_destroy(m);
        return code;
    };
    const size_t mat = (size_t)nb * nw;
    for (int r = 0; r < ndev; r++) {
        const int k0 = m->bounds[r], k1 = m->bounds[r + 1];
        const int s = wb200_create(&m->ctx[r], devices[r], nb, nw, nk, k0, k1 - k0, nn, kpb_k + (size_t)k0 * nn,
                                   bvec_cart + (size_t)k0 * nn * 3, bweights, M + 2 * (size_t)k0 * nn * nb * nb, flags);
        if (s != WB200_OK) return fail(s, std::string("wb200_multi_create: slab ") + std::to_string(r) + ": " + wb200_last_error(nullptr));
        if (cudaSetDevice(devices[r]) != cudaSuccess || wb_dev_alloc(&m->Ufull[r], sizeof(cplx) * mat * nk) != cudaSuccess)
            return fail(WB200_ERR_CUDA, "wb200_multi_create: cudaMalloc of the gauge array failed");
    }
    // peer access between the distinct devices (the pack kernels read neighbour gauges from their owners)
    for (int r = 0; r < ndev; r++)
        for (int q = 0; q < ndev; q++) {
            if (devices[r] == devices[q]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[r], devices[q]);
            if (!can) return fail(WB200_ERR_CUDA, "wb200_multi_create: no peer access between devices " +
                                                       std::to_string(devices[r]) + " and " + std::to_string(devices[q]));
            cudaSetDevice(devices[r]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[q], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(WB200_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            cudaGetLastError();
        }
    for (int p = 0; p < 2; p++)
        if (wb_pinned_alloc(&m->h_x[p], sizeof(double) * (size_t)ndev * m->maxc) != cudaSuccess ||
            wb_pinned_alloc(&m->h_r[p], sizeof(double) * (size_t)ndev * m->maxc) != cudaSuccess)
            return fail(WB200_ERR_CUDA, "wb200_multi_create: cudaMallocHost failed");
    // The library's own peer-memory collectives (peer.cu): packed halo rows pulled by the copy engines under the
    // rotation, all-reduces as a push into every slab's slot + a rank-order sum.  Between the slabs of ONE process the
    // ordering comes from CUDA events and a host rendezvous of the slab threads (WbInproc) — no kernel ever spins, so
    // repeated devices, pageable host arrays and lazily loaded kernels are all safe.
    // WB200_PEER_COMM=0: round-1 scheme (host-side all-reduce at the barrier, peer rows packed straight from peer
    // memory); =force: the spinning-flag protocol of the one-process-per-GPU mode (the test suite runs it in a
    // subprocess to cover that code on a one-GPU box).
    bool distinct = ndev <= WB_COMM_MAXR;
    bool spin_flags = false;
    {
        const char* e = getenv("WB200_PEER_COMM");
        if (e && e[0] == '0') distinct = false;
        if (e && e[0] == 'f') spin_flags = true;
    }
    const bool peer_comm = distinct && ndev > 1;
    m->flags.assign(ndev, nullptr);
    const size_t comm_bytes = wb200_peer_comm_bytes(nw);
    if (peer_comm)
        for (int r = 0; r < ndev; r++)
            if (cudaSetDevice(devices[r]) != cudaSuccess || wb_dev_alloc(&m->flags[r], comm_bytes) != cudaSuccess ||
                cudaMemset(m->flags[r], 0, comm_bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess)
                return fail(WB200_ERR_CUDA, "wb200_multi_create: cudaMalloc of the comm buffers failed");
    if (peer_comm && !spin_flags) {
        WbInproc& ip = m->inproc;
        ip.n = ndev;
        ip.barrier = [m]() { return m->bar.wait(); };
        for (auto* v : {&ip.ev_ready, &ip.ev_packed, &ip.ev_ar[0], &ip.ev_ar[1]}) {
            v->assign(ndev, nullptr);
            for (int r = 0; r < ndev; r++)
                if (cudaSetDevice(devices[r]) != cudaSuccess ||
                    cudaEventCreateWithFlags(&(*v)[r], cudaEventDisableTiming) != cudaSuccess)
                    return fail(WB200_ERR_CUDA, "wb200_multi_create: cudaEventCreate failed");
        }
    }
    std::vector<const void*> ptrs(ndev);
    std::vector<void*> packed(ndev, nullptr);
    bool have_packed = peer_comm;
    for (int r = 0; r < ndev; r++) {
        ptrs[r] = m->Ufull[r];
        size_t bytes = 0;
        if (have_packed && wb200_packed_gauge(m->ctx[r], &packed[r], &bytes) != WB200_OK) have_packed = false;
    }
    for (int r = 0; r < ndev; r++) {
        m->users[r] = {m, r};
        int s = wb200_set_peer_gauges(m->ctx[r], ndev, ptrs.data(), m->bounds.data());
        if (s == WB200_OK && peer_comm && !spin_flags) m->ctx[r]->inproc = &m->inproc;
        if (s == WB200_OK && peer_comm) s = wb200_set_peer_comm(m->ctx[r], ndev, m->flags.data());
        if (s == WB200_OK && have_packed) s = wb200_set_peer_packed(m->ctx[r], ndev, packed.data());
        if (s == WB200_OK && !peer_comm) s = wb200_set_allreduce(m->ctx[r], multi_allreduce, &m->users[r]);
        if (s != WB200_OK) return fail(s, "wb200_multi_create: " + m->ctx[r]->err);
    }
    *out = m;
    return WB200_OK;
}

extern "C" int wb200_multi_ndev(const wb200_multi* m) { return m ? m->n : -1; }
extern "C" int wb200_multi_slab_begin(const wb200_multi* m, int r) { return (m && r >= 0 && r <= m->n) ? m->bounds[r] : -1; }

extern "C" int wb200_multi_set_penalty(wb200_multi* m, const double* r0, double lambda) {
    if (!m) return WB200_ERR_ARG;
    for (auto* c : m->ctx) WB_TRY(wb200_set_penalty(c, r0, lambda));
    return WB200_OK;
}

extern "C" int wb200_multi_set_frozen(wb200_multi* m, const uint8_t* frozen) {
    if (!m) return WB200_ERR_ARG;
    for (int r = 0; r < m->n; r++)
        WB_TRY(wb200_set_frozen(m->ctx[r], frozen ? frozen + (size_t)m->bounds[r] * m->nb : nullptr));
    return WB200_OK;
}

extern "C" int wb200_multi_fg(wb200_multi* m, const double* U, double* f, double* G) {
    if (!m || !U) return WB200_ERR_ARG;
    const size_t mat2 = 2 * (size_t)m->nb * m->nw;
    std::vector<double> fr(m->n, 0.0);
    const int s = run_all(m, [&](int r) {
        const size_t off = (size_t)m->bounds[r] * mat2;
        return wb200_fg(m->ctx[r], U + off, &fr[r], G ? G + off : nullptr);
    });
    if (s == WB200_OK && f) *f = fr[0];
    return s;
}

extern "C" int wb200_multi_spread(wb200_multi* m, const double* U, double out5[5], double* omega_n, double* r,
                                  double* min_abs_Nnn) {
    if (!m || !U) return WB200_ERR_ARG;
    const size_t mat2 = 2 * (size_t)m->nb * m->nw;
    return run_all(m, [&](int q) {
        const bool lead = q == 0;   // every slab holds the same reduced result; slab 0 reports it
        return wb200_spread(m->ctx[q], U + (size_t)m->bounds[q] * mat2, lead ? out5 : nullptr, lead ? omega_n : nullptr,
                            lead ? r : nullptr, lead ? min_abs_Nnn : nullptr);
    });
}

extern "C" int wb200_multi_fg_xy(wb200_multi* m, const double* XY, double* f, double* G_XY) {
    if (!m || !XY) return WB200_ERR_ARG;
    const size_t st2 = 2 * ((size_t)m->nw * m->nw + (size_t)m->nb * m->nw);
    std::vector<double> fr(m->n, 0.0);
    const int s = run_all(m, [&](int r) {
        const size_t off = (size_t)m->bounds[r] * st2;
        return wb200_fg_xy(m->ctx[r], XY + off, &fr[r], G_XY ? G_XY + off : nullptr);
    });
    if (s == WB200_OK && f) *f = fr[0];
    return s;
}

extern "C" int wb200_multi_max_localize(wb200_multi* m, double* U_inout, double f_tol, double g_tol, int max_iter,
                                        int history, double* f_final, int* iters, int* n_fg) {
    if (!m || !U_inout) return WB200_ERR_ARG;
    const size_t mat2 = 2 * (size_t)m->nb * m->nw;
    std::vector<double> fr(m->n, 0.0);
    std::vector<int> it(m->n, 0), nf(m->n, 0);
    const int s = run_all(m, [&](int r) {
        return wb200_max_localize(m->ctx[r], U_inout + (size_t)m->bounds[r] * mat2, f_tol, g_tol, max_iter, history,
                                  &fr[r], &it[r], &nf[r]);
    });
    if (s != WB200_OK) return s;
    for (int r = 1; r < m->n; r++)
        if (it[r] != it[0] || nf[r] != nf[0] || fr[r] != fr[0]) {
            m->err = "wb200_multi_max_localize: the slabs took different paths (internal error)";
            return WB200_ERR_NOCONV;
        }
    if (f_final) *f_final = fr[0];
    if (iters) *iters = it[0];
    if (n_fg) *n_fg = nf[0];
    return WB200_OK;
}

extern "C" int wb200_multi_disentangle(wb200_multi* m, double* XY_inout, double f_tol, double g_tol, int max_iter,
                                       int history, double* f_final, int* iters, int* n_fg) {
    if (!m || !XY_inout) return WB200_ERR_ARG;
    const size_t st2 = 2 * ((size_t)m->nw * m->nw + (size_t)m->nb * m->nw);
    std::vector<double> fr(m->n, 0.0);
    std::vector<int> it(m->n, 0), nf(m->n, 0);
    const int s = run_all(m, [&](int r) {
        return wb200_disentangle(m->ctx[r], XY_inout + (size_t)m->bounds[r] * st2, f_tol, g_tol, max_iter, history,
                                 &fr[r], &it[r], &nf[r]);
    });
    if (s != WB200_OK) return s;
    for (int r = 1; r < m->n; r++)
        if (it[r] != it[0] || nf[r] != nf[0] || fr[r] != fr[0]) {
            m->err = "wb200_multi_disentangle: the slabs took different paths (internal error)";
            return WB200_ERR_NOCONV;
        }
    if (f_final) *f_final = fr[0];
    if (iters) *iters = it[0];
    if (n_fg) *n_fg = nf[0];
    return WB200_OK;
}

extern "C" long long wb200_multi_launch_count(const wb200_multi* m) {
    if (!m) return -1;
    long long t = 0;
    for (auto* c : m->ctx) t += c->launches;
    return t;
}
