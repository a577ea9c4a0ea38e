// tfhe_mgpu.cpp -- one process driving every GPU of the box (include/tfhe_b200.h, section "multi-GPU").
//
// The path shards trivially: bootstrap() reads only its own ciphertext, the shared read-only BootstrappingKey and a LUT
// (bootstrapping.rs:58-65).  So: one tfhe_ctx per device, keys replicated, contiguous balanced index ranges, and the only
// communication is the scatter of inputs and the gather of results -- per-device cudaMemcpyAsync for host buffers, grouped
// ncclSend/ncclRecv over NVLink for a batch that lives on one of the devices.  Written entirely ABOVE the single-GPU C ABI
// (no access to tfhe_ctx internals); one host thread per device because the single-GPU entry points are synchronous.
// NCCL is opened at run time (dlopen of libnccl.so.2, an already loaded copy first -- a host that links PyTorch shares its
// NCCL), so the library has no link-time dependency on it and the host-pointer form works without it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only; the symbols are resolved with dlsym
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tfhe_b200.h"

namespace {

struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string why;   // why it is unavailable
    bool ok() const { return handle != nullptr; }
};

NcclApi &nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the host process's own NCCL, if it has one
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        const char *e = dlerror();
        api.why = std::string("libnccl.so.2 cannot be opened: ") + (e ? e : "?");
        return api;
    }
#define LOAD(field, sym)                                                                     \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, sym));                          \
    if (!api.field) { api.why = std::string("libnccl lacks ") + sym; return api; }
    LOAD(CommInitAll, "ncclCommInitAll")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(Send, "ncclSend")
    LOAD(Recv, "ncclRecv")
    LOAD(GetErrorString, "ncclGetErrorString")
    LOAD(GetVersion, "ncclGetVersion")
#undef LOAD
    api.handle = h;
    return api;
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

struct tfhe_mgpu {
    tfhe_params p;
    std::vector<int> dev;
    std::vector<tfhe_ctx *> ctx;
    std::vector<ncclComm_t> comm;          // created on first device-pointer call
    std::vector<uint32_t *> stage_in[2];   // per device: staging for scattered shards (two inputs for gates)
    std::vector<uint32_t *> stage_out;
    std::vector<size_t> stage_cap;         // words, per device (same capacity for all three buffers)
    std::string err;
    double last_ms[4] = {0, 0, 0, 0};
    size_t G() const { return dev.size(); }
    size_t row() const { return (size_t)p.lwe_dimension + 1; }
};
struct tfhe_mgpu_bk {
    tfhe_mgpu *m = nullptr;
    std::vector<tfhe_bk *> bk;
};

namespace {

int mfail(tfhe_mgpu *m, int code, const std::string &msg) {
    if (m) m->err = msg;
    return code;
}
// contiguous balanced split: the first batch % G devices get one extra ciphertext
void shard(size_t total, size_t g, size_t G, size_t &lo, size_t &hi) {
    const size_t base = total / G, extra = total % G;
    lo = g * base + (g < extra ? g : extra);
    hi = lo + base + (g < extra ? 1 : 0);
}
// which of the tfhe_mgpu's devices owns p (-1: host memory, -2: a device that is not part of the tfhe_mgpu)
int owner_of(const tfhe_mgpu *m, const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return -1; }
    if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) return -1;
    for (size_t g = 0; g < m->G(); g++)
        if (m->dev[g] == a.device) return (int)g;
    return -2;
}
template <class F>
int for_each_device(tfhe_mgpu *m, F f) {   // f(g) -> status, one host thread per device; first failure wins
    std::vector<int> rc(m->G(), TFHE_OK);
    std::vector<std::thread> th;
    for (size_t g = 0; g < m->G(); g++) th.emplace_back([&, g]() { rc[g] = f(g); });
    for (auto &t : th) t.join();
    for (size_t g = 0; g < m->G(); g++)
        if (rc[g] != TFHE_OK) {
            char b[64];
            snprintf(b, sizeof b, "device %d: ", m->dev[g]);
            return mfail(m, rc[g], std::string(b) + tfhe_last_error(m->ctx[g]));
        }
    return TFHE_OK;
}
int ensure_comm(tfhe_mgpu *m) {
    if (!m->comm.empty()) return TFHE_OK;
    NcclApi &n = nccl();
    if (!n.ok()) return mfail(m, TFHE_E_NCCL, n.why);
    m->comm.assign(m->G(), nullptr);
    ncclResult_t r = n.CommInitAll(m->comm.data(), (int)m->G(), m->dev.data());
    if (r != ncclSuccess) {
        m->comm.clear();
        return mfail(m, TFHE_E_NCCL, std::string("ncclCommInitAll: ") + n.GetErrorString(r));
    }
    return TFHE_OK;
}
int ensure_stage(tfhe_mgpu *m, size_t g, size_t words) {
    if (m->stage_cap[g] >= words) return TFHE_OK;
    if (cudaSetDevice(m->dev[g]) != cudaSuccess) return mfail(m, TFHE_E_CUDA, "cudaSetDevice");
    for (std::vector<uint32_t *> *v : {&m->stage_in[0], &m->stage_in[1], &m->stage_out}) {
        if ((*v)[g]) cudaFree((*v)[g]);
        (*v)[g] = nullptr;
        if (cudaMalloc(&(*v)[g], words * 4) != cudaSuccess) { m->stage_cap[g] = 0; return mfail(m, TFHE_E_OOM, "cudaMalloc (shard staging)"); }
    }
    m->stage_cap[g] = words;
    return TFHE_OK;
}
// grouped point-to-point copies between the root and every other device: dir 0 = scatter (root -> g), 1 = gather (g -> root).
// root_base: the full [batch][row] buffer on the root; other[g]: the shard buffer on device g.
int p2p(tfhe_mgpu *m, int root, int dir, uint32_t *root_base, const std::vector<uint32_t *> &other, size_t batch) {
    NcclApi &n = nccl();
    const size_t row = m->row();
    ncclResult_t r = n.GroupStart();
    for (size_t g = 0; g < m->G() && r == ncclSuccess; g++) {
        if ((int)g == root) continue;
        size_t lo, hi;
        shard(batch, g, m->G(), lo, hi);
        if (hi == lo) continue;
        const size_t cnt = (hi - lo) * row;
        cudaStream_t sr = (cudaStream_t)tfhe_ctx_get_stream(m->ctx[root]), sg = (cudaStream_t)tfhe_ctx_get_stream(m->ctx[g]);
        if (dir == 0) {
            r = n.Send(root_base + lo * row, cnt, ncclUint32, (int)g, m->comm[root], sr);
            if (r == ncclSuccess) r = n.Recv(other[g], cnt, ncclUint32, root, m->comm[g], sg);
        } else {
            r = n.Send(other[g], cnt, ncclUint32, root, m->comm[g], sg);
            if (r == ncclSuccess) r = n.Recv(root_base + lo * row, cnt, ncclUint32, (int)g, m->comm[root], sr);
        }
    }
    ncclResult_t e = n.GroupEnd();
    if (r == ncclSuccess) r = e;
    if (r != ncclSuccess) return mfail(m, TFHE_E_NCCL, std::string(dir ? "gather: " : "scatter: ") + n.GetErrorString(r));
    return TFHE_OK;
}
int sync_all(tfhe_mgpu *m) {
    for (size_t g = 0; g < m->G(); g++) {
        if (cudaSetDevice(m->dev[g]) != cudaSuccess || cudaStreamSynchronize((cudaStream_t)tfhe_ctx_get_stream(m->ctx[g])) != cudaSuccess)
            return mfail(m, TFHE_E_CUDA, std::string("stream synchronise: ") + cudaGetErrorString(cudaGetLastError()));
    }
    return TFHE_OK;
}

// shared driver of the two batched entry points.  ins: 1 (bootstrap) or 2 (gates) input arrays; `call` runs the single-GPU
// entry point of device g on its shard [lo, hi): call(g, lo, hi, in0_shard, in1_shard, out_shard).
template <class Call>
int run_sharded(tfhe_mgpu *m, const uint32_t *const in[2], int ins, size_t batch, uint32_t *out, Call call) {
    const size_t G = m->G(), row = m->row();
    const int own_in = owner_of(m, in[0]), own_out = owner_of(m, out);
    if (own_in == -2 || own_out == -2) return mfail(m, TFHE_E_PARAM, "device pointer on a GPU that is not part of this tfhe_mgpu");
    if (own_in != own_out || (ins == 2 && owner_of(m, in[1]) != own_in))
        return mfail(m, TFHE_E_PARAM, "inputs and output must be all host memory or all on one device of the tfhe_mgpu");
    const double t0 = now_ms();
    double t1 = t0, t2;
    int rc;
    if (own_in < 0) {
        // host buffers: every device copies its own shard in and out inside its single-GPU call
        rc = for_each_device(m, [&](size_t g) {
            size_t lo, hi;
            shard(batch, g, G, lo, hi);
            return hi > lo ? call(g, lo, hi, in[0] + lo * row, ins == 2 ? in[1] + lo * row : nullptr, out + lo * row) : (int)TFHE_OK;
        });
        t2 = now_ms();
    } else {
        const int root = own_in;
        if ((rc = ensure_comm(m))) return rc;
        for (size_t g = 0; g < G; g++) {
            size_t lo, hi;
            shard(batch, g, G, lo, hi);
            if ((int)g != root && (rc = ensure_stage(m, g, (hi - lo) * row))) return rc;
        }
        for (int i = 0; i < ins; i++)
            if ((rc = p2p(m, root, 0, const_cast<uint32_t *>(in[i]), m->stage_in[i], batch))) return rc;
        if ((rc = sync_all(m))) return rc;   // (only to attribute time to the scatter; the streams would order it anyway)
        t1 = now_ms();
        rc = for_each_device(m, [&](size_t g) {
            size_t lo, hi;
            shard(batch, g, G, lo, hi);
            if (hi == lo) return (int)TFHE_OK;
            if ((int)g == root) return call(g, lo, hi, in[0] + lo * row, ins == 2 ? in[1] + lo * row : nullptr, out + lo * row);
            return call(g, lo, hi, m->stage_in[0][g], ins == 2 ? m->stage_in[1][g] : nullptr, m->stage_out[g]);
        });
        t2 = now_ms();
        if (rc == TFHE_OK) rc = p2p(m, root, 1, out, m->stage_out, batch);
        if (rc == TFHE_OK) rc = sync_all(m);
    }
    const double t3 = now_ms();
    m->last_ms[0] = t1 - t0; m->last_ms[1] = t2 - t1; m->last_ms[2] = t3 - t2; m->last_ms[3] = t3 - t0;
    return rc;
}

}  // namespace

extern "C" {

int tfhe_mgpu_create(const tfhe_params *p, int n_gpus, const int *devices, tfhe_mgpu **out) {
    if (!p || !out || n_gpus < 1) return TFHE_E_PARAM;
    if (tfhe_params_validate(p) != TFHE_OK) return TFHE_E_PARAM;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return TFHE_E_CUDA; }   // no CPU fallback
    tfhe_mgpu *m = new tfhe_mgpu();
    m->p = *p;
    for (int g = 0; g < n_gpus; g++) {
        const int d = devices ? devices[g] : g;
        bool dup = false;
        for (int e : m->dev) dup |= e == d;
        if (d < 0 || d >= count || dup) { delete m; return TFHE_E_PARAM; }
        m->dev.push_back(d);
    }
    m->ctx.assign(m->G(), nullptr);
    for (auto *v : {&m->stage_in[0], &m->stage_in[1], &m->stage_out}) v->assign(m->G(), nullptr);
    m->stage_cap.assign(m->G(), 0);
    for (size_t g = 0; g < m->G(); g++) {
        const int rc = tfhe_ctx_create(p, m->dev[g], &m->ctx[g]);
        if (rc != TFHE_OK) { tfhe_mgpu_destroy(m); return rc; }
    }
    *out = m;
    return TFHE_OK;
}

void tfhe_mgpu_destroy(tfhe_mgpu *m) {
    if (!m) return;
    for (size_t g = 0; g < m->comm.size(); g++)
        if (m->comm[g]) nccl().CommDestroy(m->comm[g]);
    for (size_t g = 0; g < m->G(); g++) {
        cudaSetDevice(m->dev[g]);
        for (auto *v : {&m->stage_in[0], &m->stage_in[1], &m->stage_out})
            if ((*v)[g]) cudaFree((*v)[g]);
        if (m->ctx[g]) tfhe_ctx_destroy(m->ctx[g]);
    }
    delete m;
}

int tfhe_mgpu_n_gpus(const tfhe_mgpu *m) { return m ? (int)m->G() : TFHE_E_PARAM; }
tfhe_ctx *tfhe_mgpu_ctx(tfhe_mgpu *m, int i) { return (m && i >= 0 && (size_t)i < m->G()) ? m->ctx[i] : nullptr; }
const char *tfhe_mgpu_last_error(const tfhe_mgpu *m) { return m ? m->err.c_str() : "null tfhe_mgpu"; }

static int upload_all(tfhe_mgpu *m, const uint32_t *bsk, const uint32_t *ksk, bool bmmp, tfhe_mgpu_bk **out) {
    if (!m || !bsk || !ksk || !out) return TFHE_E_PARAM;
    if (owner_of(m, bsk) != -1 || owner_of(m, ksk) != -1) return mfail(m, TFHE_E_PARAM, "tfhe_mgpu_bk_upload takes host pointers");
    tfhe_mgpu_bk *k = new tfhe_mgpu_bk();
    k->m = m;
    k->bk.assign(m->G(), nullptr);
    const int rc = for_each_device(m, [&](size_t g) {
        return bmmp ? tfhe_bk_upload_bmmp(m->ctx[g], bsk, ksk, &k->bk[g]) : tfhe_bk_upload(m->ctx[g], bsk, ksk, &k->bk[g]);
    });
    if (rc != TFHE_OK) { tfhe_mgpu_bk_free(k); return rc; }
    *out = k;
    return TFHE_OK;
}
int tfhe_mgpu_bk_upload(tfhe_mgpu *m, const uint32_t *bsk, const uint32_t *ksk, tfhe_mgpu_bk **out) { return upload_all(m, bsk, ksk, false, out); }
int tfhe_mgpu_bk_upload_bmmp(tfhe_mgpu *m, const uint32_t *bsk3, const uint32_t *ksk, tfhe_mgpu_bk **out) { return upload_all(m, bsk3, ksk, true, out); }
void tfhe_mgpu_bk_free(tfhe_mgpu_bk *k) {
    if (!k) return;
    for (tfhe_bk *b : k->bk) tfhe_bk_free(b);
    delete k;
}

int tfhe_mgpu_bootstrap_batch(tfhe_mgpu *m, const tfhe_mgpu_bk *bk, const uint32_t *lwe_in, const uint32_t *luts, size_t n_luts,
                              const uint32_t *lut_idx, size_t batch, uint32_t *lwe_out) {
    if (!m) return TFHE_E_PARAM;
    if (!bk || bk->m != m) return mfail(m, TFHE_E_PARAM, "bootstrapping key does not belong to this tfhe_mgpu");
    if (!lwe_in || !luts || !lwe_out || n_luts == 0) return mfail(m, TFHE_E_PARAM, "null argument");
    if (owner_of(m, luts) != -1 || (lut_idx && owner_of(m, lut_idx) != -1)) return mfail(m, TFHE_E_PARAM, "luts and lut_idx must be host pointers");
    if (batch == 0) return TFHE_OK;
    const uint32_t *in[2] = {lwe_in, nullptr};
    return run_sharded(m, in, 1, batch, lwe_out, [&](size_t g, size_t lo, size_t hi, const uint32_t *i0, const uint32_t *, uint32_t *o) {
        return tfhe_bootstrap_batch(m->ctx[g], bk->bk[g], i0, luts, n_luts, lut_idx ? lut_idx + lo : nullptr, hi - lo, o);
    });
}

int tfhe_mgpu_gates_batch(tfhe_mgpu *m, const tfhe_mgpu_bk *bk, const uint8_t *gates, const uint32_t *ct0, const uint32_t *ct1,
                          size_t batch, uint32_t *out) {
    if (!m) return TFHE_E_PARAM;
    if (!bk || bk->m != m) return mfail(m, TFHE_E_PARAM, "bootstrapping key does not belong to this tfhe_mgpu");
    if (!gates || !ct0 || !ct1 || !out) return mfail(m, TFHE_E_PARAM, "null argument");
    if (batch == 0) return TFHE_OK;
    const uint32_t *in[2] = {ct0, ct1};
    return run_sharded(m, in, 2, batch, out, [&](size_t g, size_t lo, size_t hi, const uint32_t *i0, const uint32_t *i1, uint32_t *o) {
        return tfhe_gates_batch(m->ctx[g], bk->bk[g], gates + lo, i0, i1, hi - lo, o);
    });
}

int tfhe_mgpu_last_timing(const tfhe_mgpu *m, double out[4]) {
    if (!m || !out) return TFHE_E_PARAM;
    for (int i = 0; i < 4; i++) out[i] = m->last_ms[i];
    return TFHE_OK;
}

}  // extern "C"
