// multi.cpp — one logical evaluation context over several GPUs of a box.
//
// gate_nand(ck, x, y) over a batch is embarrassingly parallel: gates are independent (gates.jl:15-18 has no shared
// mutable state) and the evaluation keys are read-only.  A tfhe_b200_multi owns one single-device context per listed
// GPU (keys replicated at load), cuts every batch into contiguous shards of whole CTA waves and runs the shards from
// one host thread per device through the ordinary host-buffer entry points, each writing a disjoint slice of the
// caller's output.  There is no data-path collective and no peer traffic; the per-device call already overlaps its
// own PCIe copies with its kernels (cabi.cu, host_chunks).
//
// Host-only translation unit: it is written against the public C ABI of the single-device context, nothing else.
#include "../../include/tfhe_b200.h"

#include <algorithm>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

struct tfhe_b200_multi {
    tfhe_b200_params P{};
    std::vector<tfhe_b200_ctx*> ctx;
    std::vector<int> dev;
    std::mutex mu;
};

namespace {

thread_local std::string g_multi_error;

int mfail(int code, const std::string& msg) {
    g_multi_error = msg;
    return code;
}

// run fn(i) on one host thread per device; first failing device wins, its message is carried to the caller's thread
int for_each_device(tfhe_b200_multi* m, const std::function<int(int)>& fn) {
    const int n = (int)m->ctx.size();
    std::vector<int> rc(n, 0);
    std::vector<std::string> msg(n);
    auto body = [&](int i) {
        rc[i] = fn(i);
        if (rc[i]) msg[i] = tfhe_b200_last_error(m->ctx[i]);   // per-thread error text: read it on the worker
    };
    std::vector<std::thread> th;
    for (int i = 1; i < n; i++) th.emplace_back(body, i);
    body(0);
    for (auto& t : th) t.join();
    for (int i = 0; i < n; i++)
        if (rc[i]) return mfail(rc[i], "device " + std::to_string(m->dev[i]) + ": " + msg[i]);
    return 0;
}

// contiguous shards: whole waves of 4 gates per SM x 148 SMs when the batch is large, an even split otherwise
void shard(size_t count, int n_dev, std::vector<size_t>& off) {
    off.assign(n_dev + 1, 0);
    const size_t wave = 4 * 148;
    size_t per = (count + n_dev - 1) / n_dev;
    if (per >= 4 * wave) per = (per + wave - 1) / wave * wave;
    for (int i = 0; i <= n_dev; i++) off[i] = std::min(count, per * (size_t)i);
    off[n_dev] = count;
}

}  // namespace

extern "C" {

int tfhe_b200_multi_create(const tfhe_b200_params* params, const int* device_ids, int n_dev, uint32_t flags,
                           tfhe_b200_multi** out) {
    if (!params || !out) return mfail(TFHE_B200_EINVAL, "null argument");
    *out = nullptr;
    const int have = tfhe_b200_device_count();
    if (have < 1) return mfail(TFHE_B200_ENODEV, "no CUDA device (this library has no CPU fallback)");
    std::vector<int> ids;
    if (!device_ids || n_dev <= 0) {
        for (int i = 0; i < have; i++) ids.push_back(i);
    } else {
        ids.assign(device_ids, device_ids + n_dev);
        for (size_t a = 0; a < ids.size(); a++) {
            if (ids[a] < 0 || ids[a] >= have) return mfail(TFHE_B200_ENODEV, "no CUDA device " + std::to_string(ids[a]));
            for (size_t b = 0; b < a; b++)
                if (ids[a] == ids[b]) return mfail(TFHE_B200_EINVAL, "device listed twice");
        }
    }
    auto* m = new tfhe_b200_multi();
    m->P = *params;
    for (int id : ids) {
        tfhe_b200_ctx* c = nullptr;
        const int rc = tfhe_b200_create(params, id, flags, &c);
        if (rc) {
            const std::string msg = tfhe_b200_last_error(nullptr);
            for (auto* k : m->ctx) tfhe_b200_destroy(k);
            delete m;
            return mfail(rc, "device " + std::to_string(id) + ": " + msg);
        }
        m->ctx.push_back(c);
        m->dev.push_back(id);
    }
    *out = m;
    return 0;
}

void tfhe_b200_multi_destroy(tfhe_b200_multi* m) {
    if (!m) return;
    for (auto* c : m->ctx) tfhe_b200_destroy(c);
    delete m;
}

const char* tfhe_b200_multi_last_error(const tfhe_b200_multi*) { return g_multi_error.c_str(); }
int tfhe_b200_multi_devices(const tfhe_b200_multi* m) { return m ? (int)m->ctx.size() : 0; }
tfhe_b200_ctx* tfhe_b200_multi_context(tfhe_b200_multi* m, int i) {
    return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr;
}
uint64_t tfhe_b200_multi_kernel_launches(const tfhe_b200_multi* m) {
    uint64_t s = 0;
    if (m) for (auto* c : m->ctx) s += tfhe_b200_kernel_launches(c);
    return s;
}

#define MULTI_LOAD(name, single)                                                     \
    int name(tfhe_b200_multi* m, const int32_t* key) {                               \
        if (!m || !key) return mfail(TFHE_B200_EINVAL, "null argument");             \
        std::lock_guard<std::mutex> lk(m->mu);                                       \
        return for_each_device(m, [&](int i) { return single(m->ctx[i], key); });    \
    }
MULTI_LOAD(tfhe_b200_multi_load_bk, tfhe_b200_load_bk)
MULTI_LOAD(tfhe_b200_multi_load_ksk, tfhe_b200_load_ksk)
MULTI_LOAD(tfhe_b200_multi_mk_load_bk, tfhe_b200_mk_load_bk)
MULTI_LOAD(tfhe_b200_multi_mk_load_ksk, tfhe_b200_mk_load_ksk)
#undef MULTI_LOAD

int tfhe_b200_multi_gate_batch(tfhe_b200_multi* m, int op, const int32_t* x, const int32_t* y, const int32_t* z,
                               int32_t* out, size_t count) {
    if (!m || !out) return mfail(TFHE_B200_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(m->mu);
    const size_t w = (size_t)m->P.n + 1;
    std::vector<size_t> off;
    shard(count, (int)m->ctx.size(), off);
    return for_each_device(m, [&](int i) {
        const size_t o = off[i], c = off[i + 1] - off[i];
        if (c == 0) return 0;
        return tfhe_b200_gate_batch(m->ctx[i], op, x ? x + o * w : nullptr, y ? y + o * w : nullptr,
                                    z ? z + o * w : nullptr, out + o * w, c);
    });
}

int tfhe_b200_multi_bootstrap_batch(tfhe_b200_multi* m, int32_t mu, const int32_t* x, int32_t* out, size_t count) {
    if (!m || !x || !out) return mfail(TFHE_B200_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(m->mu);
    const bool mk = m->P.parties > 1;
    const size_t w = (size_t)m->P.parties * m->P.n + 1;
    std::vector<size_t> off;
    shard(count, (int)m->ctx.size(), off);
    return for_each_device(m, [&](int i) {
        const size_t o = off[i], c = off[i + 1] - off[i];
        if (c == 0) return 0;
        return mk ? tfhe_b200_mk_bootstrap_batch(m->ctx[i], mu, x + o * w, out + o * w, c)
                  : tfhe_b200_bootstrap_batch(m->ctx[i], mu, x + o * w, out + o * w, c);
    });
}

int tfhe_b200_multi_mk_nand_batch(tfhe_b200_multi* m, const int32_t* x, const int32_t* y, int32_t* out, size_t count) {
    if (!m || !x || !y || !out) return mfail(TFHE_B200_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(m->mu);
    const size_t w = (size_t)m->P.parties * m->P.n + 1;
    std::vector<size_t> off;
    shard(count, (int)m->ctx.size(), off);
    return for_each_device(m, [&](int i) {
        const size_t o = off[i], c = off[i + 1] - off[i];
        if (c == 0) return 0;
        return tfhe_b200_mk_nand_batch(m->ctx[i], x + o * w, y + o * w, out + o * w, c);
    });
}

}  // extern "C"
