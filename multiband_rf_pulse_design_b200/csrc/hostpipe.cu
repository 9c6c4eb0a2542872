// see hostpipe.h
#include "hostpipe.h"

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>

namespace mbrf {
namespace hostpipe {

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int Slot::reserve_pin(size_t need)
{
    if (pin && pin_bytes >= need) return MBRF_OK;
    if (pin) { cudaFreeHost(pin); pin = nullptr; pin_bytes = 0; }
    const size_t want = need + need / 4 + 4096;
    MBRF_CUDA(cudaHostAlloc(&pin, want, cudaHostAllocPortable));
    pin_bytes = want;
    return MBRF_OK;
}

int Slot::event(size_t k, cudaEvent_t *out)
{
    while (ev.size() <= k) {
        cudaEvent_t e, e2;
        MBRF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        MBRF_CUDA(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
        ev.push_back(e);
        ev_k.push_back(e2);
    }
    *out = ev[k];
    return MBRF_OK;
}

struct Slots {
    std::map<int, std::unique_ptr<Slot>> by_device;
    ~Slots()
    {
        for (auto &kv : by_device)
            if (kv.second->pin) cudaFreeHost(kv.second->pin);   // streams / events die with the context
    }
};
static thread_local Slots t_slots;

Slot *slot(int device)
{
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(cudaGetLastError())); return nullptr; }
    auto &p = t_slots.by_device[device];
    if (!p) {
        p.reset(new Slot());
        p->device = device;
        if (cudaStreamCreateWithFlags(&p->st, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&p->st_copy, cudaStreamNonBlocking) != cudaSuccess) {
            set_error("cudaStreamCreate on device %d failed: %s", device, cudaGetErrorString(cudaGetLastError()));
            t_slots.by_device.erase(device);
            return nullptr;
        }
    }
    return p.get();
}

static int g_fanout = -1;     // -1: take MBRF_FANOUT from the environment on first use (default 1)
int fanout()
{
    if (g_fanout < 0) {
        const char *e = getenv("MBRF_FANOUT");
        int v = e ? atoi(e) : 1;
        if (e && v == 0) v = mbrf_device_count();
        g_fanout = v < 1 ? 1 : v;
    }
    return g_fanout;
}

// ---- host copy pool: the ring -> caller copies of a chunk, spread over a few threads --------------------------------------
class CopyPool {
public:
    explicit CopyPool(int workers)
    {
        for (int i = 0; i < workers; ++i) th_.emplace_back([this] { loop(); });
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> l(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    void submit(void *dst, const void *src, size_t bytes)
    {
        {
            std::lock_guard<std::mutex> l(m_);
            q_.push_back({dst, src, bytes});
            ++pending_;
        }
        cv_.notify_one();
    }
    int size() const { return (int)th_.size(); }
    void wait()      // the caller works too
    {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> l(m_);
                if (q_.empty()) {
                    done_.wait(l, [this] { return pending_ == 0; });
                    return;
                }
                j = q_.front();
                q_.pop_front();
            }
            memcpy(j.dst, j.src, j.bytes);
            finish();
        }
    }

private:
    struct Job { void *dst; const void *src; size_t bytes; };
    void finish()
    {
        std::lock_guard<std::mutex> l(m_);
        if (--pending_ == 0) done_.notify_all();
    }
    void loop()
    {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [this] { return stop_ || !q_.empty(); });
                if (stop_ && q_.empty()) return;
                j = q_.front();
                q_.pop_front();
            }
            memcpy(j.dst, j.src, j.bytes);
            finish();
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    std::deque<Job> q_;
    int pending_ = 0;
    bool stop_ = false;
};
static thread_local std::unique_ptr<CopyPool> t_pool;

static bool is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}
static double *device_alias(double *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return at.type == cudaMemoryTypeHost ? (double *)at.devicePointer : nullptr;
}

struct DevPlan {
    Slot *s = nullptr;
    long long item0 = 0, n = 0, chunk = 0;
    int nchunks = 0, ring = 0;
    size_t o_itin = 0, o_out = 0;            // offsets of the per-item-input and output rings in the pinned / device blocks
    size_t d_o_ws = 0, d_o_itin = 0, d_o_out = 0;
    size_t chunk_in_b = 0, chunk_out_b = 0;  // per component
};

int run(const Desc &d)
{
    if (d.items <= 0) return MBRF_OK;
    int cur = 0;
    MBRF_CUDA(cudaGetDevice(&cur));
    int ndev_avail = mbrf_device_count();
    int D = std::min(fanout(), ndev_avail);
    const long long min_per_dev = 65536;     // below this a second device costs more than it saves
    if ((long long)D > (d.items + min_per_dev - 1) / min_per_dev) D = (int)((d.items + min_per_dev - 1) / min_per_dev);
    if (D < 1) D = 1;
    bool pinned_out = true;
    for (int c = 0; c < d.ncomp; ++c) pinned_out = pinned_out && is_pinned(d.host_out[c]);
    const size_t item_b = d.item_doubles * 8;
    // ---- plan ----
    std::vector<DevPlan> plan((size_t)D);
    const size_t OUT_CAP = (size_t)1 << 30;  // bytes of output ring per device (device and pinned)
    bool zero_copy = false;
    for (int k = 0; k < D; ++k) {
        DevPlan &q = plan[(size_t)k];
        q.item0 = d.items * k / D;
        q.n = d.items * (k + 1) / D - q.item0;
        q.s = slot((cur + k) % ndev_avail);
        if (!q.s) return MBRF_ECUDA;
        const size_t range_b = (size_t)q.n * item_b * d.ncomp;
        // chunks: enough of them that the copy-out of one hides behind the kernels of the next, not so many that launch
        // overhead shows (every chunk re-runs the small table kernel)
        int want = (int)std::min<size_t>(8, std::max<size_t>(1, range_b / ((size_t)3 << 20)));
        if (pinned_out && D == 1 && d.item_doubles == 1 && d.ncomp_in == 0) { want = 1; zero_copy = true; }
        long long chunk = (q.n + want - 1) / want;
        chunk = (chunk + 1023) / 1024 * 1024;
        const long long cap_items = std::max<long long>(1, (long long)(OUT_CAP / 2 / (item_b * d.ncomp)));
        if (chunk > cap_items) { chunk = cap_items; zero_copy = false; }
        q.chunk = chunk;
        q.nchunks = (int)((q.n + chunk - 1) / chunk);
        q.ring = (int)std::min<long long>(q.nchunks, std::max<long long>(1, (long long)(OUT_CAP / ((size_t)chunk * item_b * d.ncomp))));
        q.chunk_out_b = align_up((size_t)chunk * item_b, 256);
        q.chunk_in_b = d.ncomp_in ? align_up((size_t)chunk * 8, 256) : 0;
        // pinned block: inputs | per-item-input ring | output ring (only when the caller's arrays are pageable)
        size_t off = align_up(d.in_bytes, 256);
        q.o_itin = off; off += (size_t)q.ring * q.chunk_in_b * d.ncomp_in;
        q.o_out = off; off += pinned_out ? 0 : (size_t)q.ring * q.chunk_out_b * d.ncomp;
        if (int rc = q.s->reserve_pin(off)) return rc;
        size_t doff = align_up(d.in_bytes, 256);
        q.d_o_ws = doff; doff += align_up(d.ws_bytes, 256);
        q.d_o_itin = doff; doff += (size_t)q.ring * q.chunk_in_b * d.ncomp_in;
        q.d_o_out = doff; doff += zero_copy ? 0 : (size_t)q.ring * q.chunk_out_b * d.ncomp;
        if (int rc = q.s->dev.reserve(doff)) return rc;
        // inputs
        d.pack((char *)q.s->pin);
        MBRF_CUDA(cudaMemcpyAsync(q.s->dev.ptr, q.s->pin, d.in_bytes, cudaMemcpyHostToDevice, q.s->st));
    }
    // host copy threads: 3 for one device, 2 more per further device (the ring -> caller copies of D devices run at once)
    const int want_threads = std::min<int>(std::max(1u, std::thread::hardware_concurrency()) - 1, 3 + 2 * (D - 1));
    if (!pinned_out && (!t_pool || t_pool->size() < want_threads)) t_pool.reset(new CopyPool(std::max(1, want_threads)));
    CopyPool *pool = pinned_out ? nullptr : t_pool.get();

    // ---- per device: waves of at most `ring` chunks -- enqueue all of them, then hand them to the copy pool in order ----
    auto work = [&](int k) -> int {
        DevPlan &q = plan[(size_t)k];
        if (cudaSetDevice(q.s->device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", q.s->device); return MBRF_ECUDA; }
        char *hp = (char *)q.s->pin, *dp = (char *)q.s->dev.ptr;
        for (int done = 0; done < q.nchunks; done += q.ring) {
            const int c1 = std::min(q.nchunks, done + q.ring);
            for (int ci = done; ci < c1; ++ci) {
                const int r = ci % q.ring;
                const long long i0 = q.item0 + (long long)ci * q.chunk, n = std::min(q.chunk, q.item0 + q.n - i0);
                const double *d_itin[4] = {nullptr, nullptr, nullptr, nullptr};
                for (int c = 0; c < d.ncomp_in; ++c) {
                    double *h = (double *)(hp + q.o_itin + ((size_t)r * d.ncomp_in + c) * q.chunk_in_b);
                    const double *src = d.host_in[c];
                    for (long long i = 0; i < n; ++i) h[i] = src[(i0 + i) * d.in_stride];
                    char *dst = dp + q.d_o_itin + ((size_t)r * d.ncomp_in + c) * q.chunk_in_b;
                    MBRF_CUDA(cudaMemcpyAsync(dst, h, (size_t)n * 8, cudaMemcpyHostToDevice, q.s->st));
                    d_itin[c] = (const double *)dst;
                }
                double *d_out[4] = {nullptr, nullptr, nullptr, nullptr};
                for (int c = 0; c < d.ncomp; ++c)
                    d_out[c] = zero_copy ? device_alias(d.host_out[c]) + (size_t)i0 * d.item_doubles
                                         : (double *)(dp + q.d_o_out + ((size_t)r * d.ncomp + c) * q.chunk_out_b);
                if (int rc = d.launch(q.s->st, dp, dp + q.d_o_ws, i0, n, d.ncomp_in ? d_itin : nullptr, d_out)) return rc;
                cudaEvent_t e;
                if (int rc = q.s->event((size_t)r, &e)) return rc;
                cudaStream_t done_on = q.s->st;
                if (!zero_copy) {
                    // results leave on the copy stream: the next chunk's kernels start while this chunk crosses PCIe
                    MBRF_CUDA(cudaEventRecord(q.s->ev_k[(size_t)r], q.s->st));
                    MBRF_CUDA(cudaStreamWaitEvent(q.s->st_copy, q.s->ev_k[(size_t)r], 0));
                    for (int c = 0; c < d.ncomp; ++c) {
                        void *dst = pinned_out ? (void *)(d.host_out[c] + (size_t)i0 * d.item_doubles)
                                               : (void *)(hp + q.o_out + ((size_t)r * d.ncomp + c) * q.chunk_out_b);
                        MBRF_CUDA(cudaMemcpyAsync(dst, d_out[c], (size_t)n * item_b, cudaMemcpyDeviceToHost, q.s->st_copy));
                    }
                    done_on = q.s->st_copy;
                }
                MBRF_CUDA(cudaEventRecord(e, done_on));
            }
            for (int ci = done; ci < c1; ++ci) {
                const int r = ci % q.ring;
                MBRF_CUDA(cudaEventSynchronize(q.s->ev[(size_t)r]));
                if (!pool) continue;
                const long long i0 = q.item0 + (long long)ci * q.chunk, n = std::min(q.chunk, q.item0 + q.n - i0);
                for (int c = 0; c < d.ncomp; ++c) {
                    const char *src = hp + q.o_out + ((size_t)r * d.ncomp + c) * q.chunk_out_b;
                    char *dst = (char *)(d.host_out[c] + (size_t)i0 * d.item_doubles);
                    const size_t bytes = (size_t)n * item_b, piece = (size_t)1 << 19;
                    for (size_t o = 0; o < bytes; o += piece) pool->submit(dst + o, src + o, std::min(piece, bytes - o));
                }
            }
            if (pool) pool->wait();      // a ring slot is reused in the next wave: its copy must be out
        }
        MBRF_CUDA(cudaStreamSynchronize(q.s->st));
        MBRF_CUDA(cudaStreamSynchronize(q.s->st_copy));
        return MBRF_OK;
    };
    std::vector<int> rcs((size_t)D, MBRF_OK);
    std::vector<std::string> errs((size_t)D);
    std::vector<std::thread> helpers;
    for (int k = 1; k < D; ++k)
        helpers.emplace_back([&, k] {
            rcs[(size_t)k] = work(k);
            if (rcs[(size_t)k]) errs[(size_t)k] = mbrf_last_error();
        });
    rcs[0] = work(0);
    for (auto &t : helpers) t.join();
    cudaSetDevice(cur);
    for (int k = 0; k < D; ++k)
        if (rcs[(size_t)k]) {
            if (k) set_error("%s", errs[(size_t)k].c_str());
            return rcs[(size_t)k];
        }
    return MBRF_OK;
}

}  // namespace hostpipe
}  // namespace mbrf

extern "C" {

int mbrf_set_fanout(int ndevices)
{
    if (ndevices < 0) return MBRF_EINVAL;
    mbrf::hostpipe::g_fanout = ndevices == 0 ? std::max(1, mbrf_device_count()) : ndevices;
    return MBRF_OK;
}

int mbrf_get_fanout(void) { return mbrf::hostpipe::fanout(); }

}  // extern "C"
