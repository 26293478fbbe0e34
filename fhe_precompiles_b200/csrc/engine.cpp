// See engine.h.
#include "engine.h"

#include <algorithm>
#include <chrono>
#include <functional>

#include "codec_kernels.h"
#include "host_pool.h"

#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <thread>

#include "kernels.h"

namespace fheb {

void cuda_throw(cudaError_t e, const char *what) {
    if (e != cudaSuccess)
        throw std::runtime_error(std::string("fhe_b200: CUDA error in ") + what + ": " + cudaGetErrorString(e));
}

namespace {
size_t env_size(const char *name, size_t dflt) {
    const char *v = getenv(name);
    if (!v || !*v) return dflt;
    long long x = atoll(v);
    return x > 0 ? (size_t)x : dflt;
}
uint64_t cheap_tag(Span s) {
    // length + 4 sampled words: a pre-filter only, equality is decided by memcmp
    uint64_t h = 0x9E3779B97F4A7C15ull ^ s.n;
    for (int i = 0; i < 4; i++) {
        size_t off = s.n >= 8 ? (s.n - 8) * (size_t)i / 3 : 0;
        uint64_t w = 0;
        memcpy(&w, s.p + off, s.n >= 8 ? 8 : s.n);
        h = (h ^ w) * 0xff51afd7ed558ccdull;
    }
    return h;
}
// Set on the workers of a batch that already runs one tile per host thread: their tiles' loops stay on the worker.  (Measured:
// with 32 tile workers on 16 cores, a nested parallel_for waited ~400 us per call for helper copies that had been started and
// then descheduled - more than the loops' own work.)
thread_local bool tl_serial_loops = false;
thread_local size_t tl_loop_width = 0;  // threads a loop of this thread may occupy, itself included (0: every core)
// fn(i) for i in [0, count): on the caller alone for short loops, otherwise shared with pool threads in chunks of `grain`
template <class F>
void parallel_for(size_t count, size_t grain, F &&fn) {
    const size_t hw = std::max(1u, std::thread::hardware_concurrency());
    if (count < 2 * grain || hw < 2 || tl_serial_loops) {
        for (size_t i = 0; i < count; i++) fn(i);
        return;
    }
    std::atomic<size_t> next{0};
    const std::function<void()> body = [&] {
        for (;;) {
            const size_t lo = next.fetch_add(grain);
            if (lo >= count) break;
            const size_t hi = std::min(count, lo + grain);
            for (size_t i = lo; i < hi; i++) fn(i);
        }
    };
    const size_t width = tl_loop_width ? std::min(tl_loop_width, hw) : hw;
    if (width < 2) {
        body();
        return;
    }
    HostPool::get().run(std::min(width - 1, (count + grain - 1) / grain - 1), body);
}
// FHE_B200_TILE_PROFILE=<n >= 1>: wall time of binary_tile's phases summed over all tile workers, printed when the process exits;
// the first n calls (warm-up: lanes, pinned buffers, key upload) are not counted
struct TileProfile {
    static constexpr int kPhases = 8;
    std::atomic<uint64_t> ns[kPhases];
    std::atomic<uint64_t> tiles{0}, calls{0}, seen{0};
    uint64_t skip = 0;
    bool on = false;
    TileProfile() {
        for (auto &v : ns) v = 0;
        const char *e = getenv("FHE_B200_TILE_PROFILE");
        on = e && atoll(e) >= 1;
        skip = on ? (uint64_t)atoll(e) : 0;
    }
    ~TileProfile() {
        if (!on || !tiles.load()) return;
        static const char *names[kPhases] = {"unpack+key", "framing", "plan", "stage", "enqueue", "device wait", "wrap", "host pass"};
        fprintf(stderr, "[fhe_b200 tile profile] %llu tiles, %llu calls; us per call (summed over workers):", (unsigned long long)tiles.load(),
                (unsigned long long)calls.load());
        for (int k = 0; k < kPhases; k++) fprintf(stderr, " %s %.1f", names[k], ns[k].load() * 1e-3 / (double)calls.load());
        fprintf(stderr, "\n");
    }
};
TileProfile g_tile_profile;
struct PhaseClock {
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    bool counted = false;
    void lap(int phase) {
        if (!counted) return;
        const auto now = std::chrono::steady_clock::now();
        g_tile_profile.ns[phase] += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(now - t).count();
        t = now;
    }
};
}  // namespace

Engine &Engine::get() {
    static Engine *e = new Engine();
    return *e;
}
void Engine::set_thread_serial_loops(bool on) { tl_serial_loops = on; }
void Engine::set_thread_loop_width(size_t threads) { tl_loop_width = threads; }

Engine::Engine() {
    HostContext::get();
    n_devices_ = device_count();
    if (n_devices_ <= 0)
        throw std::runtime_error("fhe_b200: no CUDA device visible; this library has no CPU fallback");
    size_t max_dev = env_size("FHE_B200_MAX_DEVICES", (size_t)n_devices_);
    if ((size_t)n_devices_ > max_dev) n_devices_ = (int)max_dev;
    {
        const char *v = getenv("FHE_B200_HELPER_DECODE");
        helper_decode_ = !(v && *v == '0');
        v = getenv("FHE_B200_DEVICE_CODEC");
        device_codec_ = !(v && *v == '0');
        // libzstd-written operand frames of a tile are inflated on the GPU (codec_kernels.cu, zstd_plan3.cuh): 0 never, 1 always,
        // 2 (default) when the tile brings at least FHE_B200_DEVICE_ZSTD_MIN_FRAMES of them.  fhe_b200_batch cuts a large batch
        // into tiles of zstd_tile_ops() calls so that the default applies to it; single calls and small batches inflate on the
        // host (a frame is a 16 k-step chain: ~3 ms on the GPU however few frames there are, 0.11 ms on a host core).
        v = getenv("FHE_B200_DEVICE_ZSTD");
        device_zstd_ = (v && *v == '1') ? 1 : (v && *v == '0') ? 0 : 2;
        v = getenv("FHE_B200_CALL_GRAPHS");
        call_graphs_ = !(v && *v == '0');
    }
    device_zstd_min_frames_ = env_size("FHE_B200_DEVICE_ZSTD_MIN_FRAMES", 128);
    zstd_tile_ops_ = env_size("FHE_B200_ZSTD_TILE_OPS", 128);
    zstd_tile_min_batch_ = env_size("FHE_B200_ZSTD_TILE_MIN_BATCH", 2048);
    {
        const char *v = getenv("FHE_B200_HOST_INFLATE_PCT");  // share of a tile's libzstd frames the host cores inflate meanwhile
        host_inflate_pct_ = (v && *v) ? (size_t)std::min(100, std::max(0, atoi(v))) : 0;
    }
    big_tile_set_ = getenv("FHE_B200_BIG_TILE_OPS") != nullptr;
    big_tile_ops_ = env_size("FHE_B200_BIG_TILE_OPS", 0);  // fixed big-tile size (0: never); default: engine.h big_tile_ops_for
    tile_ops_ = env_size("FHE_B200_TILE_OPS", 16);
    if (tile_ops_ < 1) tile_ops_ = 1;
    chunk_ops_ = env_size("FHE_B200_CHUNK_OPS", 4096);
    fused_ = env_size("FHE_B200_FUSED", 0) != 0;
    {
        const char *v = getenv("FHE_B200_SUBCHUNK_OPS");
        subchunk_ops_ = (v && *v) ? (size_t)atoll(v) : 0;  // opt-in L2-resident fork/join pipeline (e.g. 96); 0 = off
    }
    // byte-surface devices: FHE_B200_DEVICES="0,3" restricts single calls and batches to those GPUs (default: all).
    // Under a one-process-per-GPU launcher each rank sets it to its own device.
    if (const char *v = getenv("FHE_B200_DEVICES")) {
        for (const char *p = v; *p;) {
            char *end = nullptr;
            long d = strtol(p, &end, 10);
            if (end == p) break;
            if (d >= 0 && d < n_devices_) lane_devices_.push_back((int)d);
            p = (*end == ',') ? end + 1 : end;
        }
    }
    if (lane_devices_.empty())
        for (int d = 0; d < n_devices_; d++) lane_devices_.push_back(d);
}

// Lanes (stream + pinned staging + device buffers for one in-flight call) are created on the first byte-surface call,
// so device-resident users never pay for them.  lane_mu_ held.
void Engine::create_lanes() {
    const size_t lanes_per_dev = env_size("FHE_B200_LANES", 32);
    for (int d : lane_devices_) {
        device_context(d);
        for (size_t i = 0; i < lanes_per_dev; i++) {
            std::unique_ptr<Lane> l(new Lane());
            l->device = d;
            cuda_throw(cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking), "cudaStreamCreate");
            ensure_capacity(l.get(), 1);
            lanes_.push_back(std::move(l));
        }
    }
}

void Engine::drop_graphs(Lane *l) {
    for (auto &g : l->graphs) cudaGraphExecDestroy(g.exec);
    l->graphs.clear();
}

// (re)allocates the lane's staging for `cap` calls; the lane is held by the caller and its stream is idle
void Engine::ensure_capacity(Lane *l, size_t cap) {
    if (l->cap >= cap) return;
    drop_graphs(l);
    cudaFreeHost(l->h_a), cudaFreeHost(l->h_b), cudaFreeHost(l->h_out), cudaFreeHost(l->h_plain);
    cudaFree(l->d_a), cudaFree(l->d_b), cudaFree(l->d_out), cudaFree(l->d_plain), cudaFree(l->d_scratch);
    l->cap = 0;
    const size_t ct = cap * kCtWords * 8;
    cuda_throw(cudaMallocHost((void **)&l->h_a, ct), "cudaMallocHost");
    cuda_throw(cudaMallocHost((void **)&l->h_b, ct), "cudaMallocHost");
    cuda_throw(cudaMallocHost((void **)&l->h_out, ct), "cudaMallocHost");
    cuda_throw(cudaMallocHost((void **)&l->h_plain, cap * kN * 2), "cudaMallocHost");
    cuda_throw(cudaMalloc((void **)&l->d_a, ct), "cudaMalloc");
    cuda_throw(cudaMalloc((void **)&l->d_b, ct), "cudaMalloc");
    cuda_throw(cudaMalloc((void **)&l->d_out, ct), "cudaMalloc");
    cuda_throw(cudaMalloc((void **)&l->d_plain, cap * kN * 2), "cudaMalloc");
    cuda_throw(cudaMalloc((void **)&l->d_scratch, cap * kScratchLimbsPerOp * kN * 8), "cudaMalloc");
    l->cap = cap;
}

// device codec buffers for lane->cap calls (two ciphertext operands each)
void Engine::ensure_codec(Lane *l) {
    if (l->codec_cap >= l->cap) return;
    drop_graphs(l);
    cudaFreeHost(l->h_frames), cudaFreeHost(l->h_payloads), cudaFreeHost(l->h_outframes), cudaFreeHost(l->h_jobs), cudaFreeHost(l->h_status);
    cudaFree(l->d_frames), cudaFree(l->d_payloads), cudaFree(l->d_outframes), cudaFree(l->d_jobs), cudaFree(l->d_status), cudaFree(l->d_work);
    l->codec_cap = 0;
    const size_t cap = l->cap, ops = 2 * cap;
    cuda_throw(cudaMallocHost((void **)&l->h_frames, ops * kFrameSlotBytes), "cudaMallocHost");
    cuda_throw(cudaMallocHost((void **)&l->h_payloads, ops * kPayloadStride), "cudaMallocHost");
    cuda_throw(cudaMallocHost((void **)&l->h_outframes, cap * kPackedFrameStride), "cudaMallocHost");
    cuda_throw(cudaMallocHost((void **)&l->h_jobs, ops * sizeof(CodecJob)), "cudaMallocHost");
    cuda_throw(cudaMallocHost((void **)&l->h_status, 3 * cap * sizeof(int32_t)), "cudaMallocHost");
    cuda_throw(cudaMalloc((void **)&l->d_frames, ops * kFrameSlotBytes), "cudaMalloc");
    cuda_throw(cudaMalloc((void **)&l->d_payloads, ops * kPayloadStride), "cudaMalloc");
    cuda_throw(cudaMalloc((void **)&l->d_outframes, cap * kPackedFrameStride), "cudaMalloc");
    cuda_throw(cudaMalloc((void **)&l->d_jobs, ops * sizeof(CodecJob)), "cudaMalloc");
    cuda_throw(cudaMalloc((void **)&l->d_status, 3 * cap * sizeof(int32_t)), "cudaMalloc");
    cuda_throw(cudaMalloc(&l->d_work, ops * codec_work_bytes()), "cudaMalloc");
    if (!l->d_prefix) {
        uint8_t prefix[kCtPrefixBytes];
        canonical_ct_prefix(prefix);
        cuda_throw(cudaMalloc((void **)&l->d_prefix, kCtPrefixBytes), "cudaMalloc");
        cuda_throw(cudaMemcpy(l->d_prefix, prefix, kCtPrefixBytes, cudaMemcpyHostToDevice), "upload prefix");
    }
    l->codec_cap = cap;
}

Lane *Engine::acquire_lane() {
    std::unique_lock<std::mutex> lk(lane_mu_);
    if (lanes_.empty()) create_lanes();
    // lowest free lane of the least loaded device: a lone caller keeps reusing lane 0 (warm staging, codec buffers already grown)
    for (;;) {
        Lane *best = nullptr;
        size_t best_busy = ~(size_t)0;
        for (int d : lane_devices_) {
            size_t busy = 0;
            Lane *first_free = nullptr;
            for (auto &l : lanes_) {
                if (l->device != d) continue;
                if (l->busy) busy++;
                else if (!first_free) first_free = l.get();
            }
            if (first_free && busy < best_busy) {
                best = first_free;
                best_busy = busy;
            }
        }
        if (best) {
            best->busy = true;
            return best;
        }
        lane_cv_.wait(lk);
    }
}
void Engine::release_lane(Lane *l) {
    {
        std::lock_guard<std::mutex> lk(lane_mu_);
        l->busy = false;
    }
    lane_cv_.notify_one();
}

// ---------------------------------------------------------------- key cache
KeyPin &KeyPin::operator=(KeyPin &&o) noexcept {
    if (this != &o) {
        release();
        e_ = o.e_;
        k_ = o.k_;
        o.k_ = nullptr;
    }
    return *this;
}
void KeyPin::release() {
    if (k_) e_->unpin_key(k_);
    k_ = nullptr;
}
void Engine::unpin_key(KeyEntry *k) {
    std::lock_guard<std::mutex> lk(key_mu_);
    k->users--;
}

// Content-addressed lookup.  The byte comparison (~400 KB) and a miss's parse run OUTSIDE key_mu_: candidates are pinned
// under the lock (a pinned entry is never evicted and its `bytes` never change), compared without it.
KeyEntry *Engine::find_or_parse_key(Span pk, int32_t *rc) {
    const uint64_t tag = cheap_tag(pk);
    *rc = kOk;
    for (size_t start = 0;;) {
        KeyEntry *cand = nullptr;
        {
            std::lock_guard<std::mutex> lk(key_mu_);
            for (size_t k = start; k < keys_.size(); k++)
                if (keys_[k]->tag == tag && keys_[k]->bytes.size() == pk.n) {
                    cand = keys_[k].get();
                    cand->users++;
                    cand->last_use = ++key_clock_;
                    start = k + 1;
                    break;
                }
        }
        if (!cand) break;
        if (memcmp(cand->bytes.data(), pk.p, pk.n) == 0) return cand;  // pinned
        unpin_key(cand);
    }
    std::unique_ptr<KeyEntry> e(new KeyEntry());
    e->rk.resize(kRkWords);
    e->pk.resize(kPkWords);
    *rc = decode_public_key(pk, e->pk.data(), e->rk.data(), &e->has_relin);
    if (*rc) return nullptr;
    e->bytes.assign(pk.p, pk.p + pk.n);
    e->tag = tag;
    e->d_rk.assign((size_t)n_devices_, nullptr);
    e->d_pk.assign((size_t)n_devices_, nullptr);
    e->users = 1;
    const size_t cap = env_size("FHE_B200_KEY_CACHE", 8);

    std::lock_guard<std::mutex> lk(key_mu_);
    for (auto &o : keys_)  // another thread may have inserted the same key meanwhile
        if (o->tag == tag && o->bytes == e->bytes) {
            o->users++;
            o->last_use = ++key_clock_;
            return o.get();
        }
    if (keys_.size() >= cap) {  // evict the least recently used entry nobody is using; if all are pinned, grow
        long victim = -1;
        for (size_t i = 0; i < keys_.size(); i++)
            if (keys_[i]->users == 0 && (victim < 0 || keys_[i]->last_use < keys_[(size_t)victim]->last_use)) victim = (long)i;
        if (victim >= 0) {
            int cur = 0;
            cudaGetDevice(&cur);  // the caller has already selected its lane's device: restore it afterwards
            for (int d = 0; d < n_devices_; d++) {
                uint64_t *ptrs[2] = {keys_[(size_t)victim]->d_rk[(size_t)d], keys_[(size_t)victim]->d_pk[(size_t)d]};
                for (uint64_t *ptr : ptrs)
                    if (ptr) {
                        cudaSetDevice(d);
                        cudaFree(ptr);  // no holder left: every user synchronised its stream before unpinning
                    }
            }
            cudaSetDevice(cur);
            keys_.erase(keys_.begin() + victim);
        }
    }
    e->last_use = ++key_clock_;
    keys_.push_back(std::move(e));
    return keys_.back().get();
}

int32_t Engine::relin_key(Span pk, int device, const uint64_t **d_rk, bool need_relin, KeyPin *pin) {
    int32_t rc = kOk;
    KeyEntry *hit = find_or_parse_key(pk, &rc);  // pinned on success
    if (!hit) return rc;
    KeyPin held(this, hit);
    if (!need_relin) {
        if (d_rk) *d_rk = nullptr;
        return kOk;
    }
    if (!hit->has_relin) return kErrSunscreen;  // sunscreen: relinearization keys required but absent
    {
        std::lock_guard<std::mutex> lk(key_mu_);
        uint64_t *&slot = hit->d_rk[(size_t)device];
        if (!slot) {
            cuda_throw(cudaSetDevice(device), "cudaSetDevice");
            cuda_throw(cudaMalloc((void **)&slot, kRkWords * 8), "cudaMalloc(rk)");
            cuda_throw(cudaMemcpy(slot, hit->rk.data(), kRkWords * 8, cudaMemcpyHostToDevice), "upload rk");
        }
        *d_rk = slot;
    }
    *pin = std::move(held);
    return kOk;
}

int32_t Engine::public_key(Span pk, int device, const uint64_t **d_pk, KeyPin *pin) {
    int32_t rc = kOk;
    KeyEntry *hit = find_or_parse_key(pk, &rc);  // pinned on success
    if (!hit) return rc;
    KeyPin held(this, hit);
    {
        std::lock_guard<std::mutex> lk(key_mu_);
        uint64_t *&slot = hit->d_pk[(size_t)device];
        if (!slot) {
            cuda_throw(cudaSetDevice(device), "cudaSetDevice");
            cuda_throw(cudaMalloc((void **)&slot, kPkWords * 8), "cudaMalloc(pk)");
            cuda_throw(cudaMemcpy(slot, hit->pk.data(), kPkWords * 8, cudaMemcpyHostToDevice), "upload pk");
        }
        *d_pk = slot;
    }
    *pin = std::move(held);
    return kOk;
}

const uint64_t *Engine::network_sk(int device, Span net_pri) {
    std::lock_guard<std::mutex> lk(sk_mu_);
    if (d_net_sk_.size() < (size_t)n_devices_) d_net_sk_.assign((size_t)n_devices_, nullptr);
    uint64_t *&slot = d_net_sk_[(size_t)device];
    if (!slot) {
        std::vector<uint64_t> sk(3 * kN);
        if (decode_private_key(net_pri, sk.data()) != kOk) throw std::runtime_error("fhe_b200: embedded network private key is corrupt");
        cuda_throw(cudaSetDevice(device), "cudaSetDevice");
        cuda_throw(cudaMalloc((void **)&slot, sk.size() * 8), "cudaMalloc(sk)");
        cuda_throw(cudaMemcpy(slot, sk.data(), sk.size() * 8, cudaMemcpyHostToDevice), "upload sk");
    }
    return slot;
}

// ---------------------------------------------------------------- scratch arenas
// One arena per (device, stream): two callers on different streams of one GPU never share scratch, and work already
// enqueued on a stream keeps its arena until that stream has drained (a regrow or an eviction synchronises the owning stream
// first).  A lease pins the arena while its holder is still enqueueing; at most FHE_B200_SCRATCH_ARENAS (default 4) idle arenas
// stay allocated per device (a 4,096-op arena is 7 GB).
Engine::StreamState *Engine::lease_scratch(int device, cudaStream_t s, size_t ops) {
    std::unique_lock<std::mutex> lk(arena_mu_);
    StreamState *st = nullptr;
    for (auto &x : stream_states_)
        if (x->device == device && x->stream == s) st = x.get();
    if (!st) {
        const size_t cap = env_size("FHE_B200_SCRATCH_ARENAS", 4);
        for (;;) {  // evict idle arenas of this device, least recently used first
            size_t have = 0;
            long victim = -1;
            for (size_t i = 0; i < stream_states_.size(); i++) {
                StreamState &x = *stream_states_[i];
                if (x.device != device) continue;
                have++;
                if (x.users == 0 && (victim < 0 || x.last_use < stream_states_[(size_t)victim]->last_use)) victim = (long)i;
            }
            if (have < cap || victim < 0) break;
            StreamState &v = *stream_states_[(size_t)victim];
            cuda_throw(cudaSetDevice(device), "cudaSetDevice");
            if (cudaStreamSynchronize(v.stream) != cudaSuccess) {  // the caller may have destroyed that stream since
                cudaGetLastError();
                cuda_throw(cudaDeviceSynchronize(), "sync before scratch eviction");
            }
            free_stream_state(v);
            stream_states_.erase(stream_states_.begin() + victim);
        }
        stream_states_.emplace_back(new StreamState());
        st = stream_states_.back().get();
        st->device = device;
        st->stream = s;
    }
    // a regrow waits for other holders on the same stream (two threads driving one stream) and for the stream itself
    arena_cv_.wait(lk, [&] { return st->ops >= ops || st->users == 0; });
    if (st->ops < ops) {
        cuda_throw(cudaSetDevice(device), "cudaSetDevice");
        if (st->p) {
            cuda_throw(cudaStreamSynchronize(s), "sync before scratch regrow");
            cudaFree(st->p);
            st->p = nullptr;
            st->ops = 0;
        }
        cuda_throw(cudaMalloc((void **)&st->p, ops * kScratchLimbsPerOp * kN * 8), "cudaMalloc(scratch)");
        st->ops = ops;
    }
    st->users++;
    st->last_use = ++arena_clock_;
    return st;
}
void Engine::release_scratch(StreamState *st) {
    {
        std::lock_guard<std::mutex> lk(arena_mu_);
        st->users--;
    }
    arena_cv_.notify_all();
}
void Engine::free_stream_state(StreamState &v) {
    if (v.p) cudaFree(v.p);
    v.p = nullptr;
    ForkSet &F = v.forks;
    if (F.ready) {
        for (int i = 0; i < kForkStreams; i++) {
            if (F.scratch[i]) cudaFree(F.scratch[i]);
            cudaStreamDestroy(F.stream[i]);
            cudaEventDestroy(F.join[i]);
        }
        cudaEventDestroy(F.fork);
        F.ready = false;
    }
}
size_t Engine::set_chunk_ops(long long ops) {
    std::lock_guard<std::mutex> lk(arena_mu_);
    const size_t prev = chunk_ops_;
    if (ops > 0) chunk_ops_ = (size_t)ops;
    return prev;
}
namespace {
struct ScratchLease {
    Engine *e;
    Engine::StreamState *st;
    void (Engine::*rel)(Engine::StreamState *);
    ~ScratchLease() { (e->*rel)(st); }
};
}  // namespace

// ---------------------------------------------------------------- device-resident batched ops
// scratch layout for a chunk of c ops: tens [c][15][N] | c3 [c][6][N] | ks [c][6][N]
cudaEvent_t Engine::take_event() {
    if (!event_pool_.empty()) {
        cudaEvent_t e = event_pool_.back();
        event_pool_.pop_back();
        return e;
    }
    cudaEvent_t e;
    cuda_throw(cudaEventCreate(&e), "cudaEventCreate");
    return e;
}
void Engine::set_kernel_timing(bool on) { timing_ = on; }
void Engine::kernel_timing_report(int device, double ms[kNumTimedKernels], uint64_t launches[kNumTimedKernels]) {
    cuda_throw(cudaSetDevice(device), "cudaSetDevice");
    cuda_throw(cudaDeviceSynchronize(), "sync for timing report");
    std::lock_guard<std::mutex> tlk(timed_mu_);
    for (int k = 0; k < kNumTimedKernels; k++) ms[k] = 0, launches[k] = 0;
    for (auto &t : timed_) {
        float f = 0;
        cuda_throw(cudaEventElapsedTime(&f, t.e0, t.e1), "cudaEventElapsedTime");
        ms[t.kernel] += f;
        launches[t.kernel]++;
        event_pool_.push_back(t.e0);
        event_pool_.push_back(t.e1);
    }
    timed_.clear();
}

// kernel ids: 0 k_behz_tensor, 1 k_floor_sk, 2 k_relin_ks, 3 k_relin_finish, 4 k_ext_ntt, 5 k_tensor_intt,
//             6 k_digit_ntt, 7 k_ks_intt, 8 k_ext_conv, 9 k_ks_finish, 10 k_rk_*_ksd (key preparation), 11 k_digit_ntt_ksd,
//             12 k_ks_intt_ksd, 13 k_ks_finish_ksd, 14 k_tensor_floor_d, 15 k_ks_tail_ksd
#define TIMED(id, call, what)                                        \
    do {                                                             \
        if (timed) {                                                 \
            std::lock_guard<std::mutex> tlk(timed_mu_);              \
            TimedLaunch tl{id, take_event(), take_event()};          \
            cuda_throw(cudaEventRecord(tl.e0, s), "event record");   \
            cuda_throw(call, what);                                  \
            cuda_throw(cudaEventRecord(tl.e1, s), "event record");   \
            timed_.push_back(tl);                                    \
        } else {                                                     \
            cuda_throw(call, what);                                  \
        }                                                            \
    } while (0)

// BEHZ multiply of c ops: a, b -> m.c3 (size-3 ciphertexts)
void Engine::enqueue_mul(const uint64_t *a, const uint64_t *b, const ScratchMap &m, size_t c, cudaStream_t s, bool timed) {
    const bool dual = !fused_ && behz_mode() == 0;  // tensor product on the dual base (default) or on SEAL's 61-bit Bsk limbs
    if (fused_) {
        TIMED(0, launch_behz_tensor(a, b, m.tens, c, s), "behz_tensor");
    } else {
        if (ext_split() || dual) TIMED(8, launch_ext_conv(a, b, m.nttbuf, c, s), "ext_conv");
        TIMED(4, launch_ext_ntt(a, b, m.nttbuf, c, s), "ext_ntt");
        // three dual words one after the other in one CTA: right for chunks that fill the GPU, three times the critical path for
        // a single call or a small tile (as for the key-switch tail below)
        if (dual && (fuse_tail() & 1) && c >= 96) {
            TIMED(14, launch_tensor_floor(m.nttbuf, m.c3, c, s), "tensor_floor");
            return;
        }
        TIMED(5, launch_tensor_intt(m.nttbuf, m.tens, c, s), "tensor_intt");
    }
    TIMED(1, launch_floor_sk(m.tens, m.c3, c, s, dual), "floor_sk");
}
// relinearise c size-3 ciphertexts c3 -> out
void Engine::enqueue_relin(const uint64_t *c3, const uint64_t *rk, uint64_t *out, const ScratchMap &m, size_t c, cudaStream_t s,
                           bool timed) {
    if (fused_) {
        TIMED(2, launch_relin_ks(c3, rk, m.ks, c, s), "relin_ks");
    } else {
        if (ks_dual() && c >= 96) {
            // the whole key switch on the dual base; m.tens (15 limbs per op, dead once c3 exists) holds the lifted key
            TIMED(10, launch_rk_prepare_ksd(rk, m.tens, s), "rk_prepare_ksd");
            TIMED(11, launch_digit_ntt_ksd(c3, m.dig, c, s), "digit_ntt_ksd");
            if (fuse_tail() & 2) {
                TIMED(15, launch_ks_tail_ksd(m.dig, m.tens + 24 * kN, c3, out, c, s), "ks_tail_ksd");
                return;
            }
            TIMED(12, launch_ks_intt_ksd(m.dig, m.tens + 24 * kN, m.ks, c, s), "ks_intt_ksd");
            TIMED(13, launch_ks_finish_ksd(m.ks, c3, out, c, s), "ks_finish_ksd");
            return;
        }
        TIMED(6, launch_digit_ntt(c3, m.dig, c, s), "digit_ntt");
        // the fused tail runs the three moduli one after the other in two CTAs per op: right once 2c CTAs come close to the
        // 296 resident ones (the 128-op chunks of the host-buffer pipelines included), three times the critical path for a
        // single call or a small tile
        if (ks_finish_fused() && c >= 96) {
            // m.ks (6 limbs per op) is free on this path: it holds the lane-major copy of the key
            TIMED(9, launch_ks_finish(m.dig, rk, c3, out, c, s, m.ks), "ks_finish");
            return;
        }
        TIMED(7, launch_ks_intt(m.dig, rk, m.ks, c, s), "ks_intt");
    }
    TIMED(3, launch_relin_finish(c3, m.ks, out, c, s), "relin_finish");
}

void Engine::mul_relin_forked(StreamState *st, const uint64_t *a, const uint64_t *b, const uint64_t *rk, uint64_t *out, size_t n,
                              cudaStream_t s) {
    ForkSet &F = st->forks;  // owned by this (device, stream); the lease keeps it alive while we enqueue
    const size_t sub = subchunk_ops_;
    if (!F.ready || F.ops < sub) {
        cuda_throw(cudaStreamSynchronize(s), "sync before fork-set rebuild");
        for (int i = 0; i < kForkStreams; i++) {
            if (!F.ready) {
                cuda_throw(cudaStreamCreateWithFlags(&F.stream[i], cudaStreamNonBlocking), "cudaStreamCreate");
                cuda_throw(cudaEventCreateWithFlags(&F.join[i], cudaEventDisableTiming), "cudaEventCreate");
            }
            if (F.scratch[i]) cudaFree(F.scratch[i]);
            cuda_throw(cudaMalloc((void **)&F.scratch[i], sub * kScratchLimbsPerOp * kN * 8), "cudaMalloc(sub scratch)");
        }
        if (!F.ready) cuda_throw(cudaEventCreateWithFlags(&F.fork, cudaEventDisableTiming), "cudaEventCreate");
        F.ready = true;
        F.ops = sub;
    }
    cuda_throw(cudaEventRecord(F.fork, s), "fork record");
    for (int i = 0; i < kForkStreams; i++) cuda_throw(cudaStreamWaitEvent(F.stream[i], F.fork, 0), "fork wait");
    size_t k = 0;
    for (size_t off = 0; off < n; off += sub, k++) {
        const size_t c = n - off < sub ? n - off : sub;
        const int i = (int)(k % kForkStreams);
        ScratchMap m(F.scratch[i], sub);
        enqueue_mul(a + off * kCtWords, b + off * kCtWords, m, c, F.stream[i], timing_);
        enqueue_relin(m.c3, rk, out + off * kCtWords, m, c, F.stream[i], timing_);
    }
    for (int i = 0; i < kForkStreams; i++) {
        cuda_throw(cudaEventRecord(F.join[i], F.stream[i]), "join record");
        cuda_throw(cudaStreamWaitEvent(s, F.join[i], 0), "join wait");
    }
}

void Engine::mul_relin(int device, const uint64_t *a, const uint64_t *b, const uint64_t *rk, uint64_t *out, size_t n,
                       cudaStream_t s) {
    if (n == 0) return;
    device_context(device);
    const bool forked = subchunk_ops_ && n > subchunk_ops_;
    const size_t chunk = forked ? 0 : (chunk_ops_ < n ? chunk_ops_ : n);
    ScratchLease lease{this, lease_scratch(device, s, chunk), &Engine::release_scratch};
    if (forked) {
        mul_relin_forked(lease.st, a, b, rk, out, n, s);
        return;
    }
    ScratchMap m(lease.st->p, chunk);
    for (size_t off = 0; off < n; off += chunk) {
        const size_t c = n - off < chunk ? n - off : chunk;
        enqueue_mul(a + off * kCtWords, b + off * kCtWords, m, c, s, timing_);
        enqueue_relin(m.c3, rk, out + off * kCtWords, m, c, s, timing_);
    }
}
void Engine::multiply(int device, const uint64_t *a, const uint64_t *b, uint64_t *out3, size_t n, cudaStream_t s) {
    if (n == 0) return;
    device_context(device);
    const size_t chunk = chunk_ops_ < n ? chunk_ops_ : n;
    ScratchLease lease{this, lease_scratch(device, s, chunk), &Engine::release_scratch};
    ScratchMap m(lease.st->p, chunk);
    for (size_t off = 0; off < n; off += chunk) {
        const size_t c = n - off < chunk ? n - off : chunk;
        ScratchMap mm = m;
        mm.c3 = out3 + off * 6 * kN;
        enqueue_mul(a + off * kCtWords, b + off * kCtWords, mm, c, s, false);
    }
}
void Engine::relinearize(int device, const uint64_t *c3, const uint64_t *rk, uint64_t *out, size_t n, cudaStream_t s) {
    if (n == 0) return;
    device_context(device);
    const size_t chunk = chunk_ops_ < n ? chunk_ops_ : n;
    ScratchLease lease{this, lease_scratch(device, s, chunk), &Engine::release_scratch};
    ScratchMap m(lease.st->p, chunk);
    for (size_t off = 0; off < n; off += chunk) {
        const size_t c = n - off < chunk ? n - off : chunk;
        enqueue_relin(c3 + off * 6 * kN, rk, out + off * kCtWords, m, c, s, false);
    }
}

// ---------------------------------------------------------------- host-buffer batch (H2D | kernels | D2H overlapped)
// Chunk schedule of the host pipelines: full chunks in the middle, a short first one (the kernels start after 0.5 ms of copies
// instead of 1.3) and a short last one (less to drain) - both still >= 96 ops, the size from which the fused kernel set runs.
// A function of (off, n, chunk) only: every loop over the same batch sees the same chunks.
static size_t pipe_chunk(size_t off, size_t n, size_t chunk) {
    const size_t rem = n - off, edge = 96;
    if (chunk <= edge || rem <= edge) return rem < chunk ? rem : chunk;
    if (off == 0) return edge;
    if (rem <= chunk) return rem > 2 * edge ? rem - edge : rem;
    if (rem <= chunk + edge) return rem - edge;
    return chunk;
}

void Engine::mul_relin_host(int device, const uint64_t *a, const uint64_t *b, const uint64_t *rk, uint64_t *out, size_t n) {
    device_context(device);
    {
        std::lock_guard<std::mutex> lk(arena_mu_);
        if (pipes_.size() < (size_t)n_devices_) pipes_.resize((size_t)n_devices_);
        if (!pipes_[(size_t)device]) pipes_[(size_t)device].reset(new HostPipe());
    }
    HostPipe &P = *pipes_[(size_t)device];
    std::lock_guard<std::mutex> lk(P.mu);
    if (!P.chunk) P.chunk = env_size("FHE_B200_PIPE_CHUNK_OPS", 256);  // (sweep r2w: frames 282 k ops/s at 256 against 260 k at 128; limb arrays equal within noise)
    if (!P.d_rk) cuda_throw(cudaMalloc((void **)&P.d_rk, kRkWords * 8), "cudaMalloc(rk)");
    for (auto &sl : P.slot) {
        if (!sl.stream) cuda_throw(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking), "cudaStreamCreate");
        if (!sl.d_a) {
            cuda_throw(cudaMalloc((void **)&sl.d_a, P.chunk * kCtWords * 8), "cudaMalloc");
            cuda_throw(cudaMalloc((void **)&sl.d_b, P.chunk * kCtWords * 8), "cudaMalloc");
            cuda_throw(cudaMalloc((void **)&sl.d_out, P.chunk * kCtWords * 8), "cudaMalloc");
            cuda_throw(cudaMalloc((void **)&sl.d_scratch, P.chunk * kScratchLimbsPerOp * kN * 8), "cudaMalloc");
        }
    }
    P.ready = true;
    cuda_throw(cudaMemcpyAsync(P.d_rk, rk, kRkWords * 8, cudaMemcpyHostToDevice, P.slot[0].stream), "H2D rk");
    cuda_throw(cudaStreamSynchronize(P.slot[0].stream), "sync rk");
    size_t i = 0;
    for (size_t off = 0, c = 0; off < n; off += c, i++) {
        c = pipe_chunk(off, n, P.chunk);
        PipeSlot &sl = P.slot[i % kPipeSlots];
        cudaStream_t s = sl.stream;
        cuda_throw(cudaMemcpyAsync(sl.d_a, a + off * kCtWords, c * kCtWords * 8, cudaMemcpyHostToDevice, s), "H2D a");
        cuda_throw(cudaMemcpyAsync(sl.d_b, b + off * kCtWords, c * kCtWords * 8, cudaMemcpyHostToDevice, s), "H2D b");
        ScratchMap m(sl.d_scratch, P.chunk);
        enqueue_mul(sl.d_a, sl.d_b, m, c, s, false);
        enqueue_relin(m.c3, P.d_rk, sl.d_out, m, c, s, false);
        cuda_throw(cudaMemcpyAsync(out + off * kCtWords, sl.d_out, c * kCtWords * 8, cudaMemcpyDeviceToHost, s), "D2H out");
    }
    for (auto &sl : P.slot) cuda_throw(cudaStreamSynchronize(sl.stream), "pipe sync");
}

// ---------------------------------------------------------------- serialized operands (structured frames), pipelined like the above
// Operand i of an array is the zstd frame at f + i * stride (the body of a ciphertext this library serialized: 82,054 bytes,
// codec_kernels.h).  Frames are copied as they are (5 bytes per residue instead of 8 over PCIe), validated and unpacked on the
// GPU, and the result is written as a frame on the GPU: fout + i * kPackedFrameStride.  status[i]: 0 done, 1 an operand is not a
// structured frame (or fails the range checks) - result undefined, use the byte surface -, 2 the result needs the generic writer.
void Engine::mul_relin_frames(int device, const uint8_t *fa, const uint8_t *fb, size_t stride, const uint64_t *rk, uint8_t *fout,
                              size_t n, int32_t *status) {
    if (stride < kPackedFrameBytes) throw std::runtime_error("mul_relin_frames: stride smaller than a structured frame");
    device_context(device);
    {
        std::lock_guard<std::mutex> lk(arena_mu_);
        if (pipes_.size() < (size_t)n_devices_) pipes_.resize((size_t)n_devices_);
        if (!pipes_[(size_t)device]) pipes_[(size_t)device].reset(new HostPipe());
    }
    HostPipe &P = *pipes_[(size_t)device];
    std::lock_guard<std::mutex> lk(P.mu);
    if (!P.chunk) P.chunk = env_size("FHE_B200_PIPE_CHUNK_OPS", 256);  // (sweep r2w: frames 282 k ops/s at 256 against 260 k at 128; limb arrays equal within noise)
    if (!P.d_rk) cuda_throw(cudaMalloc((void **)&P.d_rk, kRkWords * 8), "cudaMalloc(rk)");
    if (!P.d_prefix) {
        uint8_t prefix[kCtPrefixBytes];
        canonical_ct_prefix(prefix);
        cuda_throw(cudaMalloc((void **)&P.d_prefix, kCtPrefixBytes), "cudaMalloc");
        cuda_throw(cudaMemcpy(P.d_prefix, prefix, kCtPrefixBytes, cudaMemcpyHostToDevice), "upload prefix");
    }
    const size_t region = P.chunk * stride;  // one operand array of a chunk
    for (auto &sl : P.slot) {
        if (!sl.stream) cuda_throw(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking), "cudaStreamCreate");
        if (!sl.d_a) {
            cuda_throw(cudaMalloc((void **)&sl.d_a, P.chunk * kCtWords * 8), "cudaMalloc");
            cuda_throw(cudaMalloc((void **)&sl.d_b, P.chunk * kCtWords * 8), "cudaMalloc");
            cuda_throw(cudaMalloc((void **)&sl.d_out, P.chunk * kCtWords * 8), "cudaMalloc");
            cuda_throw(cudaMalloc((void **)&sl.d_scratch, P.chunk * kScratchLimbsPerOp * kN * 8), "cudaMalloc");
        }
        if (sl.frame_stride != stride) {  // staging [pad | a frames | pad | b frames | pad], jobs interleaved a0 b0 a1 b1 ...
            cudaFree(sl.d_frames), cudaFree(sl.d_outframes), cudaFree(sl.d_jobs), cudaFree(sl.d_status);
            cuda_throw(cudaMalloc((void **)&sl.d_frames, 2 * region + 3 * kFramePad), "cudaMalloc");
            cuda_throw(cudaMemset(sl.d_frames, 0, 2 * region + 3 * kFramePad), "cudaMemset");
            cuda_throw(cudaMalloc((void **)&sl.d_outframes, P.chunk * kPackedFrameStride), "cudaMalloc");
            cuda_throw(cudaMalloc((void **)&sl.d_jobs, 2 * P.chunk * sizeof(CodecJob)), "cudaMalloc");
            cuda_throw(cudaMalloc((void **)&sl.d_status, 3 * P.chunk * sizeof(int32_t)), "cudaMalloc");
            std::vector<CodecJob> jobs(2 * P.chunk);
            for (size_t i = 0; i < P.chunk; i++) {
                jobs[2 * i] = CodecJob{kFramePad + i * stride, (uint32_t)kPackedFrameBytes, kJobPacked, (int32_t)i, 0};
                jobs[2 * i + 1] = CodecJob{2 * kFramePad + region + i * stride, (uint32_t)kPackedFrameBytes, kJobPacked, (int32_t)i, 1};
            }
            cuda_throw(cudaMemcpy(sl.d_jobs, jobs.data(), jobs.size() * sizeof(CodecJob), cudaMemcpyHostToDevice), "upload jobs");
            sl.frame_stride = stride;
        }
    }
    P.ready = true;
    if (P.h_status_cap < 3 * n) {
        if (P.h_status) cudaFreeHost(P.h_status);
        cuda_throw(cudaMallocHost((void **)&P.h_status, 3 * n * sizeof(int32_t)), "cudaMallocHost");
        P.h_status_cap = 3 * n;
    }
    cuda_throw(cudaMemcpyAsync(P.d_rk, rk, kRkWords * 8, cudaMemcpyHostToDevice, P.slot[0].stream), "H2D rk");
    cuda_throw(cudaStreamSynchronize(P.slot[0].stream), "sync rk");
    size_t k = 0;
    for (size_t off = 0, c = 0; off < n; off += c, k++) {
        c = pipe_chunk(off, n, P.chunk);
        PipeSlot &sl = P.slot[k % kPipeSlots];
        cudaStream_t s = sl.stream;
        // the last frame of an array may end before its stride does: never read past fa / fb + (n-1) * stride + frame
        const size_t bytes = (c - 1) * stride + kPackedFrameBytes;
        cuda_throw(cudaMemcpyAsync(sl.d_frames + kFramePad, fa + off * stride, bytes, cudaMemcpyHostToDevice, s), "H2D frames a");
        cuda_throw(cudaMemcpyAsync(sl.d_frames + 2 * kFramePad + region, fb + off * stride, bytes, cudaMemcpyHostToDevice, s), "H2D frames b");
        cuda_throw(launch_codec_inflate(sl.d_frames, nullptr, sl.d_jobs, sl.d_status, nullptr, P.d_prefix, sl.d_a, sl.d_b, (int)(2 * c), false,
                                        true, false, s),
                   "frame unpack");
        ScratchMap m(sl.d_scratch, P.chunk);
        enqueue_mul(sl.d_a, sl.d_b, m, c, s, false);
        enqueue_relin(m.c3, P.d_rk, sl.d_out, m, c, s, false);
        cuda_throw(launch_codec_pack(sl.d_out, sl.d_outframes, sl.d_status + 2 * P.chunk, P.d_prefix, (int)c, s), "frame pack");
        cuda_throw(cudaMemcpyAsync(fout + off * kPackedFrameStride, sl.d_outframes, (c - 1) * kPackedFrameStride + kPackedFrameBytes,
                                   cudaMemcpyDeviceToHost, s),
                   "D2H frames");
        cuda_throw(cudaMemcpyAsync(P.h_status + 3 * off, sl.d_status, 2 * c * sizeof(int32_t), cudaMemcpyDeviceToHost, s), "D2H status");
        cuda_throw(cudaMemcpyAsync(P.h_status + 3 * off + 2 * c, sl.d_status + 2 * P.chunk, c * sizeof(int32_t), cudaMemcpyDeviceToHost, s),
                   "D2H flags");
    }
    for (auto &sl : P.slot) cuda_throw(cudaStreamSynchronize(sl.stream), "pipe sync");
    if (status) {
        for (size_t off = 0, c = 0; off < n; off += c) {
            c = pipe_chunk(off, n, P.chunk);
            const int32_t *st = P.h_status + 3 * off;
            for (size_t i = 0; i < c; i++) {
                const bool ok = st[2 * i] != kJobFallback && st[2 * i + 1] != kJobFallback;  // the unpack kernel only flags failures
                status[off + i] = !ok ? 1 : (st[2 * c + i] ? 2 : 0);
            }
        }
    }
}

// ---------------------------------------------------------------- serialized <-> device-resident ciphertexts
Engine::XferPipe &Engine::xfer_pipe(int device, size_t n, size_t slot_bytes) {
    {
        std::lock_guard<std::mutex> lk(arena_mu_);
        if (xfers_.size() < (size_t)n_devices_) xfers_.resize((size_t)n_devices_);
        if (!xfers_[(size_t)device]) xfers_[(size_t)device].reset(new XferPipe());
    }
    XferPipe &X = *xfers_[(size_t)device];
    if (!X.d_prefix) {
        uint8_t prefix[kCtPrefixBytes];
        canonical_ct_prefix(prefix);
        cuda_throw(cudaMalloc((void **)&X.d_prefix, kCtPrefixBytes), "cudaMalloc");
        cuda_throw(cudaMemcpy(X.d_prefix, prefix, kCtPrefixBytes, cudaMemcpyHostToDevice), "upload prefix");
    }
    for (auto &sl : X.slot) {
        if (!sl.stream) {
            cuda_throw(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking), "cudaStreamCreate");
            cuda_throw(cudaMalloc((void **)&sl.d_jobs, kXferChunk * sizeof(CodecJob)), "cudaMalloc");
            cuda_throw(cudaMallocHost((void **)&sl.h_jobs, kXferChunk * sizeof(CodecJob)), "cudaMallocHost");
            cuda_throw(cudaMalloc((void **)&sl.d_status, kXferChunk * sizeof(int32_t)), "cudaMalloc");
        }
        if (sl.d_frames_bytes < slot_bytes) {
            cuda_throw(cudaStreamSynchronize(sl.stream), "sync before staging regrow");
            cudaFree(sl.d_frames);
            cuda_throw(cudaMalloc((void **)&sl.d_frames, slot_bytes), "cudaMalloc");
            cuda_throw(cudaMemset(sl.d_frames, 0, slot_bytes), "cudaMemset");
            sl.d_frames_bytes = slot_bytes;
        }
    }
    if (X.h_status_cap < n) {
        if (X.h_status) cudaFreeHost(X.h_status);
        cuda_throw(cudaMallocHost((void **)&X.h_status, n * sizeof(int32_t)), "cudaMallocHost");
        X.h_status_cap = n;
    }
    return X;
}

// frames (host, `stride` apart) -> d_words [n][2][2][N] on `device`.  status[i]: 0 ok, 1 not a structured frame or a residue out
// of range (the words of that ciphertext are then undefined).
void Engine::upload_frames(int device, const uint8_t *frames, size_t stride, size_t n, uint64_t *d_words, int32_t *status) {
    if (n == 0) return;
    if (stride < kPackedFrameBytes) throw std::runtime_error("upload_frames: stride smaller than a structured frame");
    device_context(device);
    XferPipe &X = xfer_pipe(device, n, kXferChunk * stride + 2 * kFramePad);
    std::lock_guard<std::mutex> lk(X.mu);
    size_t k = 0;
    for (size_t off = 0; off < n; off += kXferChunk, k++) {
        const size_t c = std::min(kXferChunk, n - off);
        XferSlot &sl = X.slot[k & 1];
        cudaStream_t s = sl.stream;
        if (k >= 2) cuda_throw(cudaStreamSynchronize(s), "slot sync");  // h_jobs of this slot are free again
        for (size_t i = 0; i < c; i++) sl.h_jobs[i] = CodecJob{kFramePad + i * stride, (uint32_t)kPackedFrameBytes, kJobPacked, (int32_t)i, 0};
        cuda_throw(cudaMemcpyAsync(sl.d_jobs, sl.h_jobs, c * sizeof(CodecJob), cudaMemcpyHostToDevice, s), "H2D jobs");
        cuda_throw(cudaMemcpyAsync(sl.d_frames + kFramePad, frames + off * stride, (c - 1) * stride + kPackedFrameBytes, cudaMemcpyHostToDevice, s),
                   "H2D frames");
        cuda_throw(launch_codec_inflate(sl.d_frames, nullptr, sl.d_jobs, sl.d_status, nullptr, X.d_prefix, d_words + off * kCtWords, nullptr,
                                        (int)c, false, true, false, s),
                   "frame unpack");
        cuda_throw(cudaMemcpyAsync(X.h_status + off, sl.d_status, c * sizeof(int32_t), cudaMemcpyDeviceToHost, s), "D2H status");
    }
    for (auto &sl : X.slot) cuda_throw(cudaStreamSynchronize(sl.stream), "xfer sync");
    if (status)
        for (size_t i = 0; i < n; i++) status[i] = X.h_status[i] == kJobFallback ? 1 : 0;
}

// d_words [n][2][2][N] on `device` -> n structured frames (host, kPackedFrameStride apart).  status[i]: 0 ok, 2 the ciphertext is
// constant (transparent) and has no structured frame: serialise it with fhe_b200_write_ciphertext.
void Engine::download_frames(int device, const uint64_t *d_words, size_t n, uint8_t *out_frames, int32_t *status) {
    if (n == 0) return;
    device_context(device);
    XferPipe &X = xfer_pipe(device, n, kXferChunk * kPackedFrameStride);
    std::lock_guard<std::mutex> lk(X.mu);
    size_t k = 0;
    for (size_t off = 0; off < n; off += kXferChunk, k++) {
        const size_t c = std::min(kXferChunk, n - off);
        XferSlot &sl = X.slot[k & 1];
        cudaStream_t s = sl.stream;
        cuda_throw(launch_codec_pack(d_words + off * kCtWords, sl.d_frames, sl.d_status, X.d_prefix, (int)c, s), "frame pack");
        cuda_throw(cudaMemcpyAsync(out_frames + off * kPackedFrameStride, sl.d_frames, (c - 1) * kPackedFrameStride + kPackedFrameBytes,
                                   cudaMemcpyDeviceToHost, s),
                   "D2H frames");
        cuda_throw(cudaMemcpyAsync(X.h_status + off, sl.d_status, c * sizeof(int32_t), cudaMemcpyDeviceToHost, s), "D2H flags");
    }
    for (auto &sl : X.slot) cuda_throw(cudaStreamSynchronize(sl.stream), "xfer sync");
    if (status)
        for (size_t i = 0; i < n; i++) status[i] = X.h_status[i] ? 2 : 0;
}

// ---------------------------------------------------------------- byte surface, one call
namespace {
struct LaneGuard {
    Engine *e;
    Lane *l;
    void (Engine::*rel)(Lane *);
    ~LaneGuard() { (e->*rel)(l); }
};
}  // namespace

namespace {
thread_local CallBreakdown tl_breakdown;
inline double us_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
}
}  // namespace
const CallBreakdown &Engine::last_call_breakdown() { return tl_breakdown; }

int32_t Engine::binary_op(Op op, Shape shape, Kind kind, Span in, std::vector<uint8_t> *out) {
    TileItem it{op, shape, kind, in, {}, 0};
    binary_tile(&it, 1);
    if (it.rc == 0) out->swap(it.out);
    return it.rc;
}

// One call whose inputs are perfectly ordinary (the overwhelmingly common case), arranged for latency: the public-key
// comparison and the inflate of the first operand run on the caller while a pool thread inflates the second operand; payloads
// are validated and unpacked on the GPU and the result frame is written there.  Returns false -- having produced nothing --
// whenever anything deviates; binary_tile's general path then redoes the call and owns every error code.
bool Engine::single_call_fast(Lane *lane, TileItem &it, bool timed, std::chrono::steady_clock::time_point t_start) {
    Span pk, sa, sb;
    if (unpack_binary_operation(it.in, &pk, &sa, &sb)) return false;
    const int nct = it.shape == Shape::CtCt ? 2 : 1;
    const Span cts[2] = {it.shape == Shape::PtCt ? sb : sa, sb};
    CipherView views[2];
    Span frames[2];
    int kinds[2] = {0, 0};
    for (int k = 0; k < nct; k++) {
        Span blob;
        uint8_t compr = 0;
        if (parse_ciphertext_framing(cts[k], &views[k], &blob) != kOk || !data_type_matches(views[k].data_type, it.kind)) return false;
        kinds[k] = classify_ciphertext_blob(blob, &frames[k], &compr);
        if (kinds[k] < 1 || frames[k].n + 2 * kFramePad > kFrameSlotBytes) return false;
        views[k].compr_mode = compr;
    }
    if (zstd_writer() != 1) return false;
    ensure_codec(lane);
    cudaStream_t s = lane->stream;

    // operand k is job k: payload slot k (libzstd frames, inflated here) or a structured frame copied as it is
    const size_t frame_off[2] = {kFramePad, kFramePad + kFrameSlotBytes};  // fixed places: the copies below are replayed as a graph
    const auto stage = [&](int k) -> bool {
        if (kinds[k] == 2) {
            memcpy(lane->h_frames + frame_off[k], frames[k].p, frames[k].n);
            return true;
        }
        return inflate_ct_payload(frames[k], lane->h_payloads + (size_t)k * kPayloadStride);
    };
    const bool need_relin = it.op == Op::Mul && it.shape == Shape::CtCt;
    const uint64_t *d_rk = nullptr;
    KeyPin pin;
    int32_t key_rc = kOk;
    bool staged[2] = {true, true};
    std::atomic<int> next{0};
    const std::function<void()> work = [&] {  // tasks: first operand, second operand (ct x ct only), key comparison
        for (int t; (t = next.fetch_add(1)) < 3;) {
            if (t == 2) key_rc = relin_key(pk, lane->device, &d_rk, need_relin, &pin);
            else if (t < nct) staged[t] = stage(t);
        }
    };
    HostPool::get().run((size_t)nct, work);
    if (key_rc != kOk || !staged[0] || !staged[1]) return false;
    if (it.shape != Shape::CtCt && encode_scalar(it.kind, it.shape == Shape::CtPt ? sb : sa, lane->h_plain) != kOk) return false;
    const double t_decode = us_since(t_start);

    bool any_payload = false, any_packed = false;
    for (int k = 0; k < nct; k++) {
        lane->h_jobs[k] = CodecJob{frame_off[k], (uint32_t)frames[k].n, kinds[k] == 2 ? kJobPacked : kJobPayload, 0, k};
        (kinds[k] == 2 ? any_packed : any_payload) = true;
    }
    int32_t *d_cflag = lane->d_status + nct, *h_cflag = lane->h_status + nct;  // right behind the job status: one copy back
    const auto enqueue = [&](bool ev) {
        if (ev) cudaEventRecord(lane->ev[0], s);
        cuda_throw(cudaMemcpyAsync(lane->d_jobs, lane->h_jobs, (size_t)nct * sizeof(CodecJob), cudaMemcpyHostToDevice, s), "H2D jobs");
        for (int k = 0; k < nct; k++)
            if (kinds[k] == 2)  // a structured frame has one size
                cuda_throw(cudaMemcpyAsync(lane->d_frames + frame_off[k] - kFramePad, lane->h_frames + frame_off[k] - kFramePad,
                                           kPackedFrameBytes + 2 * kFramePad, cudaMemcpyHostToDevice, s),
                           "H2D frame");
        if (any_payload) {  // slots 0 and 1 are adjacent: one copy covers both when both are payloads
            const int first = kinds[0] == 2 ? 1 : 0, last = (nct == 2 && kinds[1] != 2) ? 1 : 0;
            if (last >= first)
                cuda_throw(cudaMemcpyAsync(lane->d_payloads + (size_t)first * kPayloadStride, lane->h_payloads + (size_t)first * kPayloadStride,
                                           (size_t)(last - first + 1) * kPayloadStride, cudaMemcpyHostToDevice, s),
                           "H2D payloads");
        }
        if (it.shape != Shape::CtCt) cuda_throw(cudaMemcpyAsync(lane->d_plain, lane->h_plain, kN * 2, cudaMemcpyHostToDevice, s), "H2D plain");
        if (ev) cudaEventRecord(lane->ev[1], s);
        cuda_throw(launch_codec_inflate(lane->d_frames, lane->d_payloads, lane->d_jobs, lane->d_status, lane->d_work, lane->d_prefix, lane->d_a,
                                        lane->d_b, nct, false, any_packed, any_payload, s),
                   "codec unpack");
        if (it.shape == Shape::CtCt) {
            if (it.op == Op::Mul) {
                ScratchMap m(lane->d_scratch, 1);
                enqueue_mul(lane->d_a, lane->d_b, m, 1, s, false);
                enqueue_relin(m.c3, d_rk, lane->d_out, m, 1, s, false);
            } else {
                cuda_throw(launch_eltwise(lane->d_a, lane->d_b, lane->d_out, 1, it.op == Op::Add ? 0 : 1, s), "eltwise");
            }
        } else if (it.op == Op::Mul) {
            cuda_throw(launch_mul_plain(lane->d_a, lane->d_plain, lane->d_out, 1, s), "mul_plain");
        } else {
            const int mode = it.op == Op::Add ? 0 : (it.shape == Shape::CtPt ? 1 : 3);
            cuda_throw(launch_plain_addsub(lane->d_a, lane->d_plain, lane->d_out, 1, mode, s), "plain_addsub");
        }
        cuda_throw(launch_codec_pack(lane->d_out, lane->d_outframes, d_cflag, lane->d_prefix, 1, s), "codec pack");
        if (ev) cudaEventRecord(lane->ev[2], s);
        cuda_throw(cudaMemcpyAsync(lane->h_outframes, lane->d_outframes, kPackedFrameBytes, cudaMemcpyDeviceToHost, s), "D2H frame");
        cuda_throw(cudaMemcpyAsync(lane->h_status, lane->d_status, (size_t)(nct + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, s), "D2H status");
        if (ev) cudaEventRecord(lane->ev[3], s);
    };
    // every size and pointer above is fixed for a given call shape on this lane (payload slots, structured frames and the
    // result frame have constant sizes), so the sequence is replayed as a CUDA graph: ~13 launches and copies become one
    if (call_graphs_ && !timed) {
        cudaGraphExec_t exec = nullptr;
        for (const auto &g : lane->graphs)
            if (g.op == (int)it.op && g.shape == (int)it.shape && g.kind0 == kinds[0] && g.kind1 == kinds[1] && g.rk == d_rk) {
                exec = g.exec;
                count_launches(g.launches);
            }
        if (!exec) {
            const uint64_t before = launch_count() + codec_launch_count();
            if (lane->graphs.size() >= 16) drop_graphs(lane);
            cudaGraph_t graph = nullptr;
            cuda_throw(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal), "begin capture");
            try {
                enqueue(false);
            } catch (...) {
                cudaStreamEndCapture(s, &graph);
                if (graph) cudaGraphDestroy(graph);
                throw;
            }
            cuda_throw(cudaStreamEndCapture(s, &graph), "end capture");
            cuda_throw(cudaGraphInstantiate(&exec, graph, 0), "graph instantiate");
            cudaGraphDestroy(graph);
            lane->graphs.push_back(Lane::CallGraph{(int)it.op, (int)it.shape, kinds[0], kinds[1], d_rk, exec,
                                                   launch_count() + codec_launch_count() - before});
        }
        cuda_throw(cudaGraphLaunch(exec, s), "graph launch");
    } else {
        enqueue(timed);
    }
    cuda_throw(cudaStreamSynchronize(s), "stream sync");
    const double t_device = us_since(t_start);
    for (int k = 0; k < nct; k++)
        if (lane->h_status[k] == kJobFallback) return false;
    if (h_cflag[0]) {  // constant result (a transparent ciphertext): the host writer's libzstd path
        cuda_throw(cudaMemcpyAsync(lane->h_out, lane->d_out, kCtWords * 8, cudaMemcpyDeviceToHost, s), "D2H");
        cuda_throw(cudaStreamSynchronize(s), "stream sync");
        it.rc = encode_ciphertext(views[0], lane->h_out, &it.out);
    } else {
        wrap_ciphertext_blob(views[0], lane->h_outframes, kPackedFrameBytes, &it.out);
        it.rc = kOk;
    }
    if (timed) {
        CallBreakdown &b = tl_breakdown;
        float ms[3] = {0, 0, 0};
        for (int k = 0; k < 3; k++) cudaEventElapsedTime(&ms[k], lane->ev[k], lane->ev[k + 1]);
        b.unpack_key_us = 0;  // the key comparison overlaps the inflates on this path
        b.decode_us = t_decode;
        b.h2d_us = ms[0] * 1e3;
        b.kernels_us = ms[1] * 1e3;
        b.d2h_us = ms[2] * 1e3;
        b.total_us = us_since(t_start);
        b.encode_us = b.total_us - t_device;
    }
    return true;
}

void Engine::binary_tile(TileItem *items, size_t cnt) {
    if (cnt == 0) return;
    const bool timed = call_timing_.load(std::memory_order_relaxed);
    const auto t_start = std::chrono::steady_clock::now();
    Lane *lane = acquire_lane();
    LaneGuard guard{this, lane, &Engine::release_lane};
    cuda_throw(cudaSetDevice(lane->device), "cudaSetDevice");
    ensure_capacity(lane, cnt);
    cudaStream_t s = lane->stream;
    if (timed && !lane->ev[0])
        for (auto &e : lane->ev) cuda_throw(cudaEventCreate(&e), "cudaEventCreate");

    if (cnt == 1 && device_codec_ && helper_decode_ && single_call_fast(lane, items[0], timed, t_start)) return;

    // device work classes; ct-ct first so that the second operand array is one prefix of the slots
    enum Cls { kMulCt = 0, kAddCt, kSubCt, kMulPt, kAddPt, kSubCtPt, kSubPtCt };
    struct Prep {
        Span sa, sb;
        const uint64_t *d_rk = nullptr;
        KeyPin pin;
        int32_t key_rc = 0;
        CipherView va;
        int cls = 0;
        bool live = false;
    };
    std::vector<Prep> prep(cnt);
    std::vector<size_t> order;
    order.reserve(cnt);
    PhaseClock pc;
    if (g_tile_profile.on && g_tile_profile.seen.fetch_add(cnt) >= g_tile_profile.skip)
        pc.counted = true, g_tile_profile.tiles++, g_tile_profile.calls += cnt;

    // pass 1, reference order (pack.rs:261-263): framing, then the public key (a ~400 KB comparison with the cached key per call:
    // shared with pool threads when the tile is large)
    parallel_for(cnt, 4, [&](size_t i) {
        TileItem &it = items[i];
        Prep &p = prep[i];
        Span pk;
        if ((it.rc = unpack_binary_operation(it.in, &pk, &p.sa, &p.sb))) return;
        const bool need_relin = (it.op == Op::Mul && it.shape == Shape::CtCt);
        int32_t rc = relin_key(pk, lane->device, &p.d_rk, need_relin, &p.pin);
        if (rc == kErrSunscreen && need_relin) {
            // missing relin keys is a runtime (not a decoding) error: operands are still decoded first
        } else if (rc) {
            it.rc = rc;
            return;
        }
        p.key_rc = rc;
        if (it.shape == Shape::CtCt) p.cls = it.op == Op::Mul ? kMulCt : (it.op == Op::Add ? kAddCt : kSubCt);
        else if (it.op == Op::Mul) p.cls = kMulPt;
        else if (it.op == Op::Add) p.cls = kAddPt;
        else p.cls = it.shape == Shape::CtPt ? kSubCtPt : kSubPtCt;
        p.live = true;
    });
    for (size_t i = 0; i < cnt; i++)
        if (prep[i].live) order.push_back(i);
    pc.lap(0);
    const double t_unpack = us_since(t_start);
    struct Run {
        int cls;
        const uint64_t *d_rk;
        size_t begin, end;
    };
    const auto by_class = [&](size_t x, size_t y) {
        if (prep[x].cls != prep[y].cls) return prep[x].cls < prep[y].cls;
        return prep[x].cls == kMulCt && prep[x].d_rk < prep[y].d_rk;
    };
    const auto launch_runs = [&](const std::vector<Run> &rs) {
        for (const Run &r : rs) {
            const size_t c = r.end - r.begin;
            const uint64_t *a = lane->d_a + r.begin * kCtWords;
            const uint64_t *b = lane->d_b + r.begin * kCtWords;
            const uint16_t *pl = lane->d_plain + r.begin * kN;
            uint64_t *o = lane->d_out + r.begin * kCtWords;
            switch (r.cls) {
                case kMulCt: {
                    ScratchMap m(lane->d_scratch, c);
                    enqueue_mul(a, b, m, c, s, false);
                    enqueue_relin(m.c3, r.d_rk, o, m, c, s, false);
                    break;
                }
                case kAddCt: cuda_throw(launch_eltwise(a, b, o, c, 0, s), "eltwise"); break;
                case kSubCt: cuda_throw(launch_eltwise(a, b, o, c, 1, s), "eltwise"); break;
                case kMulPt: cuda_throw(launch_mul_plain(a, pl, o, c, s), "mul_plain"); break;
                // a + b: add_plain; ct - pt: sub_plain; pt - ct: negate(sub_plain(ct, pt))  
                case kAddPt: cuda_throw(launch_plain_addsub(a, pl, o, c, 0, s), "plain_addsub"); break;
                case kSubCtPt: cuda_throw(launch_plain_addsub(a, pl, o, c, 1, s), "plain_addsub"); break;
                default: cuda_throw(launch_plain_addsub(a, pl, o, c, 3, s), "plain_addsub"); break;
            }
        }
    };
    std::stable_sort(order.begin(), order.end(), by_class);

    // ---- device codec pass (tiles only): calls whose operands are zstd-mode ciphertexts with clean framing travel as
    // compressed frames; the GPU inflates, validates, computes and packs the result frame.  Whatever is not perfectly
    // ordinary -- and every operand the strict device decoder hands back -- is left for the host pass below, which
    // reproduces the reference's checks and error codes in their order.
    if (device_codec_ && cnt >= 2 && !order.empty()) {
        ensure_codec(lane);
        std::vector<Run> runs;
        std::vector<size_t> slot_item, rest;
        const bool pack_on_device = zstd_writer() == 1;
        // (A) per call, in parallel: bincode framing, type check, classification of the SEAL blobs
        struct Stg {
            bool clean = false;
            int nct = 0, kinds[2] = {0, 0};
            Span frames[2];
            CipherView views[2];
            bool on_host[2] = {false, false};  // libzstd-written frame inflated on the host (else staged for the device decoder)
            int job0 = 0;
            size_t slot = 0, off[2] = {0, 0};
            std::atomic<bool> bad{false};      // staging failed: the host pass redoes the call
        };
        std::vector<Stg> stg(order.size());
        parallel_for(order.size(), 8, [&](size_t oi) {
            const size_t i = order[oi];
            TileItem &it = items[i];
            Prep &p = prep[i];
            Stg &g = stg[oi];
            bool clean = p.key_rc == 0;
            const Span cts[2] = {it.shape == Shape::PtCt ? p.sb : p.sa, p.sb};
            const int want_ct = it.shape == Shape::CtCt ? 2 : 1;
            for (int k = 0; clean && k < want_ct; k++) {
                Span blob;
                uint8_t compr = 0;
                clean = parse_ciphertext_framing(cts[k], &g.views[k], &blob) == kOk && data_type_matches(g.views[k].data_type, it.kind);
                if (!clean) break;
                g.kinds[k] = classify_ciphertext_blob(blob, &g.frames[k], &compr);
                clean = g.kinds[k] >= 1 && g.frames[k].n + 2 * kFramePad <= kFrameSlotBytes;
                g.views[k].compr_mode = compr;
                g.nct++;
            }
            g.clean = clean;
        });
        pc.lap(1);
        // (B) serial: slots, jobs and staging offsets.  libzstd-written frames go to the device decoder when it is on -- always
        // (FHE_B200_DEVICE_ZSTD=1) or, by default, when the tile brings enough frames to fill the GPU -- except for the share
        // the host cores inflate meanwhile (FHE_B200_HOST_INFLATE_PCT): both decoders work on the same tile at once.
        size_t lib_frames = 0;
        for (const Stg &g : stg)
            if (g.clean)
                for (int k = 0; k < g.nct; k++) lib_frames += g.kinds[k] == 1;
        const bool dev_zstd = device_zstd_ == 1 || (device_zstd_ == 2 && lib_frames >= device_zstd_min_frames_);
        size_t slots = 0, ctct_slots = 0, fcur = 0, seen_lib = 0;
        int njobs = 0;
        bool any_zstd = false, any_packed = false, any_payload = false;
        std::vector<size_t> slot_stg;
        for (size_t oi = 0; oi < order.size(); oi++) {
            const size_t i = order[oi];
            Stg &g = stg[oi];
            Prep &p = prep[i];
            if (!g.clean) {
                rest.push_back(i);
                continue;
            }
            p.va = g.views[0];
            g.slot = slots;
            g.job0 = njobs;
            for (int k = 0; k < g.nct; k++) {
                bool host = g.kinds[k] == 1 && !dev_zstd;
                if (g.kinds[k] == 1 && dev_zstd) {  // Bresenham split of the libzstd frames between host and device
                    host = (seen_lib + 1) * host_inflate_pct_ / 100 != seen_lib * host_inflate_pct_ / 100;
                    seen_lib++;
                }
                g.on_host[k] = host;
                if (host) {
                    lane->h_jobs[njobs++] = CodecJob{0, 0, kJobPayload, (int32_t)slots, k};  // payload slot = job index
                    any_payload = true;
                    continue;
                }
                g.off[k] = fcur + kFramePad;
                lane->h_jobs[njobs++] = CodecJob{g.off[k], (uint32_t)g.frames[k].n, g.kinds[k] == 2 ? kJobPacked : kJobZstd, (int32_t)slots, k};
                (g.kinds[k] == 2 ? any_packed : any_zstd) = true;
                fcur += (g.frames[k].n + 2 * kFramePad + 15) & ~(size_t)15;
            }
            const uint64_t *rk = p.cls == kMulCt ? p.d_rk : nullptr;
            if (runs.empty() || runs.back().cls != p.cls || runs.back().d_rk != rk) runs.push_back(Run{p.cls, rk, slots, slots});
            runs.back().end = ++slots;
            if (p.cls <= kSubCt) ctct_slots = slots;
            slot_item.push_back(i);
            slot_stg.push_back(oi);
        }
        pc.lap(2);
        if (slots) {
            // (C) bytes into the pinned staging -- frames as they are, encoded scalars -- and the device decoder is started on the
            // frames; (C2) the payloads the host inflates follow while the GPU decodes (hybrid tiles: FHE_B200_HOST_INFLATE_PCT)
            const auto fail_slot = [&](Stg &g) {
                g.bad = true;
                for (int c = 0; c < g.nct; c++) lane->h_jobs[g.job0 + c].kind = kJobNone;  // the kernels skip it; its slot computes on stale data
            };
            parallel_for(slots, 2, [&](size_t k) {
                Stg &g = stg[slot_stg[k]];
                TileItem &it = items[slot_item[k]];
                Prep &p = prep[slot_item[k]];
                for (int c = 0; c < g.nct; c++)
                    if (!g.on_host[c]) memcpy(lane->h_frames + g.off[c], g.frames[c].p, g.frames[c].n);
                if (it.shape != Shape::CtCt && encode_scalar(it.kind, it.shape == Shape::CtPt ? p.sb : p.sa, lane->h_plain + k * kN) != kOk)
                    fail_slot(g);
            });
            pc.lap(3);
            const bool split_codec = any_zstd && any_payload;  // decode on the device first, inflate the host's share meanwhile
            if (njobs) {
                if (fcur) cuda_throw(cudaMemcpyAsync(lane->d_frames, lane->h_frames, fcur, cudaMemcpyHostToDevice, s), "H2D frames");
                cuda_throw(cudaMemcpyAsync(lane->d_jobs, lane->h_jobs, (size_t)njobs * sizeof(CodecJob), cudaMemcpyHostToDevice, s), "H2D jobs");
            }
            if (split_codec)
                cuda_throw(launch_codec_inflate(lane->d_frames, lane->d_payloads, lane->d_jobs, lane->d_status, lane->d_work, lane->d_prefix,
                                                lane->d_a, lane->d_b, njobs, any_zstd, any_packed, any_payload, s, kCodecDecode),
                           "codec decode");
            if (any_payload) {
                std::atomic<bool> any_failed{false};
                parallel_for(slots, 2, [&](size_t k) {
                    Stg &g = stg[slot_stg[k]];
                    if (g.bad) return;
                    for (int c = 0; c < g.nct; c++)
                        if (g.on_host[c] && !inflate_ct_payload(g.frames[c], lane->h_payloads + (size_t)(g.job0 + c) * kPayloadStride)) {
                            fail_slot(g);
                            any_failed = true;
                            break;
                        }
                });
                if (any_failed)  // (the first copy of the job list may or may not have seen the cleared kinds: either is fine)
                    cuda_throw(cudaMemcpyAsync(lane->d_jobs, lane->h_jobs, (size_t)njobs * sizeof(CodecJob), cudaMemcpyHostToDevice, s), "H2D jobs");
            }
            // payloads inflated on the host: one copy per run of adjacent payload slots
            for (int j0 = 0; j0 < njobs;) {
                if (lane->h_jobs[j0].kind != kJobPayload) {
                    j0++;
                    continue;
                }
                int j1 = j0 + 1;
                while (j1 < njobs && lane->h_jobs[j1].kind == kJobPayload) j1++;
                cuda_throw(cudaMemcpyAsync(lane->d_payloads + (size_t)j0 * kPayloadStride, lane->h_payloads + (size_t)j0 * kPayloadStride,
                                           (size_t)(j1 - j0) * kPayloadStride, cudaMemcpyHostToDevice, s),
                           "H2D payloads");
                j0 = j1;
            }
            if (slots > ctct_slots)
                cuda_throw(cudaMemcpyAsync(lane->d_plain + ctct_slots * kN, lane->h_plain + ctct_slots * kN,
                                           (slots - ctct_slots) * kN * 2, cudaMemcpyHostToDevice, s),
                           "H2D plain");
            cuda_throw(launch_codec_inflate(lane->d_frames, lane->d_payloads, lane->d_jobs, lane->d_status, lane->d_work, lane->d_prefix,
                                            lane->d_a, lane->d_b, njobs, any_zstd, any_packed, any_payload, s,
                                            split_codec ? kCodecUnpack : kCodecAll),
                       "codec inflate");
            launch_runs(runs);
            int32_t *d_cflag = lane->d_status + 2 * lane->cap, *h_cflag = lane->h_status + 2 * lane->cap;
            if (pack_on_device) {
                cuda_throw(launch_codec_pack(lane->d_out, lane->d_outframes, d_cflag, lane->d_prefix, (int)slots, s), "codec pack");
                cuda_throw(cudaMemcpyAsync(lane->h_outframes, lane->d_outframes, slots * kPackedFrameStride, cudaMemcpyDeviceToHost, s),
                           "D2H frames");
                cuda_throw(cudaMemcpyAsync(h_cflag, d_cflag, slots * sizeof(int32_t), cudaMemcpyDeviceToHost, s), "D2H flags");
            } else {
                cuda_throw(cudaMemcpyAsync(lane->h_out, lane->d_out, slots * kCtWords * 8, cudaMemcpyDeviceToHost, s), "D2H");
            }
            if (njobs)
                cuda_throw(cudaMemcpyAsync(lane->h_status, lane->d_status, (size_t)njobs * sizeof(int32_t), cudaMemcpyDeviceToHost, s),
                           "D2H status");
            pc.lap(4);
            cuda_throw(cudaStreamSynchronize(s), "stream sync");
            pc.lap(5);
            // (D) in parallel: results into their own buffers; calls the device handed back are collected for the host pass
            std::vector<uint8_t> redo(slots, 0);
            parallel_for(slots, 4, [&](size_t k) {
                const size_t i = slot_item[k];
                Stg &g = stg[slot_stg[k]];
                bool ok = !g.bad;
                for (int j = g.job0; ok && j < g.job0 + g.nct; j++) ok = lane->h_status[j] != kJobFallback;
                if (!ok) {
                    redo[k] = 1;
                    return;
                }
                TileItem &it = items[i];
                if (pack_on_device && !h_cflag[k]) {
                    const uint8_t *frame = lane->h_outframes + k * kPackedFrameStride;
                    uint8_t *direct = nullptr;
                    if (it.take_malloc) {
                        it.out_malloc_len = wrapped_ciphertext_size(prep[i].va, kPackedFrameBytes);
                        direct = (uint8_t *)malloc(it.out_malloc_len);
                    }
                    if (direct) {
                        wrap_ciphertext_blob_to(prep[i].va, frame, kPackedFrameBytes, direct);
                        it.out_malloc = direct;
                    } else {
                        wrap_ciphertext_blob(prep[i].va, frame, kPackedFrameBytes, &it.out);
                    }
                    it.rc = kOk;
                } else if (!pack_on_device) {
                    it.rc = encode_ciphertext(prep[i].va, lane->h_out + k * kCtWords, &it.out);
                } else {
                    redo[k] = 2;  // constant result (a transparent ciphertext): the host writer's libzstd path, below
                }
                if (redo[k] == 0) prep[i].live = false;
            });
            for (size_t k = 0; k < slots; k++) {
                const size_t i = slot_item[k];
                if (redo[k] == 1) {
                    rest.push_back(i);
                } else if (redo[k] == 2) {
                    cuda_throw(cudaMemcpyAsync(lane->h_out + k * kCtWords, lane->d_out + k * kCtWords, kCtWords * 8, cudaMemcpyDeviceToHost, s),
                               "D2H");
                    cuda_throw(cudaStreamSynchronize(s), "stream sync");
                    items[i].rc = encode_ciphertext(prep[i].va, lane->h_out + k * kCtWords, &items[i].out);
                    prep[i].live = false;
                }
            }
        }
        std::stable_sort(rest.begin(), rest.end(), by_class);
        order.swap(rest);
        pc.lap(6);
    }

    // ---- host pass: operands into adjacent staging slots, class by class; a call that fails here gives its slot to the next
    std::vector<Run> runs;
    std::vector<size_t> slot_item;
    size_t slots = 0, ctct_slots = 0;
    for (size_t i : order) {
        TileItem &it = items[i];
        Prep &p = prep[i];
        uint64_t *ha = lane->h_a + slots * kCtWords;
        uint64_t *hb = lane->h_b + slots * kCtWords;
        uint16_t *hp = lane->h_plain + slots * kN;
        int32_t rc = 0;
        if (it.shape == Shape::CtCt) {
            CipherView vb;
            if (cnt == 1 && helper_decode_) {
                // a single call: inflate the second operand on a pool thread while this one inflates the first (a tile
                // already keeps every core busy).  If no pool thread is free the caller does both, in order.
                int32_t rcs[2] = {0, 0};
                std::atomic<int> next{0};
                const std::function<void()> fn = [&] {
                    for (int k; (k = next.fetch_add(1)) < 2;)
                        rcs[k] = k == 0 ? decode_ciphertext(p.sa, &p.va, ha) : decode_ciphertext(p.sb, &vb, hb);
                };
                HostPool::get().run(1, fn);
                rc = rcs[0] ? rcs[0] : rcs[1];  // reference order: a's error wins
            } else if (!(rc = decode_ciphertext(p.sa, &p.va, ha))) {
                rc = decode_ciphertext(p.sb, &vb, hb);
            }
            if (!rc && (!data_type_matches(p.va.data_type, it.kind) || !data_type_matches(vb.data_type, it.kind))) rc = kErrSunscreen;
        } else if (it.shape == Shape::CtPt) {
            if (!(rc = decode_ciphertext(p.sa, &p.va, ha)) && !(rc = encode_scalar(it.kind, p.sb, hp)))
                if (!data_type_matches(p.va.data_type, it.kind)) rc = kErrSunscreen;
        } else {
            if (!(rc = encode_scalar(it.kind, p.sa, hp)) && !(rc = decode_ciphertext(p.sb, &p.va, ha)))
                if (!data_type_matches(p.va.data_type, it.kind)) rc = kErrSunscreen;
        }
        if (!rc) rc = p.key_rc;
        if (rc) {
            it.rc = rc;
            p.live = false;
            continue;
        }
        const uint64_t *rk = p.cls == kMulCt ? p.d_rk : nullptr;
        if (runs.empty() || runs.back().cls != p.cls || runs.back().d_rk != rk) runs.push_back(Run{p.cls, rk, slots, slots});
        runs.back().end = ++slots;
        if (p.cls <= kSubCt) ctct_slots = slots;
        slot_item.push_back(i);
    }
    if (slots == 0) return;
    const double t_decode = us_since(t_start);

    if (timed) cudaEventRecord(lane->ev[0], s);
    cuda_throw(cudaMemcpyAsync(lane->d_a, lane->h_a, slots * kCtWords * 8, cudaMemcpyHostToDevice, s), "H2D a");
    if (ctct_slots)
        cuda_throw(cudaMemcpyAsync(lane->d_b, lane->h_b, ctct_slots * kCtWords * 8, cudaMemcpyHostToDevice, s), "H2D b");
    if (slots > ctct_slots)
        cuda_throw(cudaMemcpyAsync(lane->d_plain + ctct_slots * kN, lane->h_plain + ctct_slots * kN, (slots - ctct_slots) * kN * 2,
                                   cudaMemcpyHostToDevice, s),
                   "H2D plain");
    if (timed) cudaEventRecord(lane->ev[1], s);
    launch_runs(runs);
    if (timed) cudaEventRecord(lane->ev[2], s);
    cuda_throw(cudaMemcpyAsync(lane->h_out, lane->d_out, slots * kCtWords * 8, cudaMemcpyDeviceToHost, s), "D2H");
    if (timed) cudaEventRecord(lane->ev[3], s);
    cuda_throw(cudaStreamSynchronize(s), "stream sync");
    const double t_device = us_since(t_start);
    for (size_t k = 0; k < slots; k++) {
        TileItem &it = items[slot_item[k]];
        it.rc = encode_ciphertext(prep[slot_item[k]].va, lane->h_out + k * kCtWords, &it.out);
    }
    pc.lap(7);
    if (timed) {
        CallBreakdown &b = tl_breakdown;
        float ms[3] = {0, 0, 0};
        for (int k = 0; k < 3; k++) cudaEventElapsedTime(&ms[k], lane->ev[k], lane->ev[k + 1]);
        b.unpack_key_us = t_unpack;
        b.decode_us = t_decode - t_unpack;
        b.h2d_us = ms[0] * 1e3;
        b.kernels_us = ms[1] * 1e3;
        b.d2h_us = ms[2] * 1e3;
        b.total_us = us_since(t_start);
        b.encode_us = b.total_us - t_device;
    }
}

// ---------------------------------------------------------------- device-resident encrypt / decrypt
void Engine::encrypt_device(int device, const uint64_t *pk, const uint16_t *plain, const uint64_t *seeds, uint64_t *ct, size_t n,
                            cudaStream_t s) {
    device_context(device);
    if (n == 0) return;
    const size_t chunk = chunk_ops_ < n ? chunk_ops_ : n;
    ScratchLease lease{this, lease_scratch(device, s, chunk), &Engine::release_scratch};
    uint64_t *encbuf = lease.st->p;
    for (size_t off = 0; off < n; off += chunk) {
        const size_t c = n - off < chunk ? n - off : chunk;
        cuda_throw(launch_encrypt(pk, plain + off * kN, seeds + off * 8, encbuf, ct + off * kCtWords, c, s), "encrypt");
    }
}
void Engine::decrypt_device(int device, const uint64_t *ct, const uint64_t *sk, uint16_t *plain, size_t n, cudaStream_t s,
                            int32_t *exhausted) {
    device_context(device);
    if (n == 0) return;
    const size_t chunk = chunk_ops_ < n ? chunk_ops_ : n;
    ScratchLease lease{this, lease_scratch(device, s, chunk), &Engine::release_scratch};
    uint64_t *xbuf = lease.st->p;
    for (size_t off = 0; off < n; off += chunk) {
        const size_t c = n - off < chunk ? n - off : chunk;
        cuda_throw(launch_decrypt(ct + off * kCtWords, sk, xbuf, plain + off * kN, c, s, exhausted ? exhausted + off : nullptr), "decrypt");
    }
}

// ---------------------------------------------------------------- threshold-network simulation API (fhe.rs:594-779)
namespace {
// the reference's private 512-bit constant mixed into the encrypt seed (fhe.rs:604-609)
const uint8_t kSeedConstant[64] = {15,  17,  225, 5,   30,  1,   237, 218, 130, 19,  37,  95,  222, 218, 244, 172,
                                   214, 175, 175, 110, 173, 103, 172, 60,  43,  76,  40,  150, 215, 96,  23,  78,
                                   22,  39,  30,  177, 107, 130, 124, 109, 27,  96,  206, 125, 104, 241, 10,  40,
                                   88,  238, 117, 118, 79,  113, 213, 110, 148, 179, 53,  19,  227, 154, 151, 122};
// the eight little-endian u64 words of SHA-512(msg) the reference hands to SEAL as the PRNG seed (fhe.rs:47-54, 611-616):
// all 512 bits reach the device sampler
struct Seed512 {
    uint64_t w[8];
};
Seed512 seed_from_hash(const std::vector<uint8_t> &msg) {
    uint8_t h[64];
    sha512(msg.data(), msg.size(), h);
    Seed512 s;
    memcpy(s.w, h, 64);
    return s;
}
void testnet_params(uint8_t out[kParamsBytes]) {
    uint64_t w[6] = {(uint64_t)kN, 3, kModulus[MQ0], kModulus[MQ1], kModulus[MP], kT};
    memcpy(out, w, 48);
    memset(out + 48, 0, 8);
}
}  // namespace

// encrypts the scalar operand `scalar` under `pk_bytes` with `seed`; uses the lane's plain / out buffers
int32_t Engine::encrypt_plain(Kind kind, Span scalar, Span pk_bytes, const uint64_t seed[8], CipherView *view, Lane *lane,
                              std::vector<uint8_t> *out) {
    int32_t rc = encode_scalar(kind, scalar, lane->h_plain);
    if (rc) return rc == kErrSunscreen ? kErrFailedEncryption : rc;
    const uint64_t *d_pk = nullptr;
    KeyPin pin;
    rc = public_key(pk_bytes, lane->device, &d_pk, &pin);
    if (rc) return rc;
    cudaStream_t s = lane->stream;
    memcpy(lane->h_b, seed, 64);
    cuda_throw(cudaMemcpyAsync(lane->d_plain, lane->h_plain, kN * 2, cudaMemcpyHostToDevice, s), "H2D plain");
    cuda_throw(cudaMemcpyAsync(lane->d_b, lane->h_b, 64, cudaMemcpyHostToDevice, s), "H2D seed");
    int *d_failed = reinterpret_cast<int *>(lane->d_scratch + 6 * kN);  // right behind the [6][N] encryption scratch
    cuda_throw(launch_encrypt(d_pk, lane->d_plain, lane->d_b, lane->d_scratch, lane->d_out, 1, s, d_failed), "encrypt");
    cuda_throw(cudaMemcpyAsync(lane->h_out, lane->d_out, kCtWords * 8, cudaMemcpyDeviceToHost, s), "D2H");
    cuda_throw(cudaMemcpyAsync(lane->h_b + 8, d_failed, sizeof(int), cudaMemcpyDeviceToHost, s), "D2H sampler flag");
    cuda_throw(cudaStreamSynchronize(s), "stream sync");
    if (*reinterpret_cast<const int *>(lane->h_b + 8)) return kErrFailedEncryption;  // PRNG stream window exhausted (~24 sigma)
    return encode_ciphertext(*view, lane->h_out, out);
}

int32_t Engine::encrypt(Kind kind, Span in, Span net_pub, Span, std::vector<uint8_t> *out) {
    Span plain, public_data;
    int32_t rc = unpack_two_arguments(in, &plain, &public_data);
    if (rc) return rc;
    // seed = SHA-512(public_data || constant || plain bytes)   (fhe.rs:600-611)
    std::vector<uint8_t> msg(public_data.p, public_data.p + public_data.n);
    msg.insert(msg.end(), kSeedConstant, kSeedConstant + 64);
    msg.insert(msg.end(), plain.p, plain.p + plain.n);
    Lane *lane = acquire_lane();
    LaneGuard guard{this, lane, &Engine::release_lane};
    cuda_throw(cudaSetDevice(lane->device), "cudaSetDevice");
    CipherView view;
    view.data_type = data_type_of(kind);
    testnet_params(view.params);
    return encrypt_plain(kind, plain, net_pub, seed_from_hash(msg).w, &view, lane, out);
}

int32_t Engine::decrypt(Kind kind, Span in, Span, Span net_pri, std::vector<uint8_t> *out) {
    Lane *lane = acquire_lane();
    LaneGuard guard{this, lane, &Engine::release_lane};
    cuda_throw(cudaSetDevice(lane->device), "cudaSetDevice");
    CipherView view;
    int32_t rc = decode_ciphertext(in, &view, lane->h_a);
    if (rc) return rc == kErrSunscreen ? kErrFailedDecryption : rc;
    if (!data_type_matches(view.data_type, kind)) return kErrFailedDecryption;  // fhe.rs:696 map_err
    const uint64_t *d_sk = network_sk(lane->device, net_pri);
    cudaStream_t s = lane->stream;
    cuda_throw(cudaMemcpyAsync(lane->d_a, lane->h_a, kCtWords * 8, cudaMemcpyHostToDevice, s), "H2D ct");
    int *d_flag = reinterpret_cast<int *>(lane->d_scratch + 2 * kN);  // right behind the [2][N] dot-product buffer
    cuda_throw(launch_decrypt(lane->d_a, d_sk, lane->d_scratch, lane->d_plain, 1, s, d_flag), "decrypt");
    cuda_throw(cudaMemcpyAsync(lane->h_plain, lane->d_plain, kN * 2, cudaMemcpyDeviceToHost, s), "D2H plain");
    cuda_throw(cudaMemcpyAsync(lane->h_out, d_flag, sizeof(int), cudaMemcpyDeviceToHost, s), "D2H noise flag");
    cuda_throw(cudaStreamSynchronize(s), "stream sync");
    // sunscreen's Runtime::decrypt refuses a ciphertext whose invariant noise budget is 0 (-> fhe.rs:643, 696)
    if (*reinterpret_cast<const int *>(lane->h_out)) return kErrFailedDecryption;
    decode_scalar(kind, lane->h_plain, kN, out);
    return kOk;
}

int32_t Engine::reencrypt(Kind kind, Span in, Span, Span net_pri, std::vector<uint8_t> *out) {
    Span pk, ct, public_data;
    int32_t rc = unpack_binary_operation(in, &pk, &ct, &public_data);
    if (rc) return rc;
    Lane *lane = acquire_lane();
    LaneGuard guard{this, lane, &Engine::release_lane};
    cuda_throw(cudaSetDevice(lane->device), "cudaSetDevice");
    // reference order: key, ciphertext, public data are all deserialised first (pack.rs:261-263)
    const uint64_t *d_pk = nullptr;
    KeyPin pin;  // validates the target key first, as the reference's unpack order does; encrypt_plain re-pins it
    if ((rc = public_key(pk, lane->device, &d_pk, &pin))) return rc;
    CipherView view;
    if ((rc = decode_ciphertext(ct, &view, lane->h_a))) return rc == kErrSunscreen ? kErrFailedDecryption : rc;
    if (!data_type_matches(view.data_type, kind)) return kErrFailedDecryption;
    const uint64_t *d_sk = network_sk(lane->device, net_pri);
    cudaStream_t s = lane->stream;
    cuda_throw(cudaMemcpyAsync(lane->d_a, lane->h_a, kCtWords * 8, cudaMemcpyHostToDevice, s), "H2D ct");
    int *d_flag = reinterpret_cast<int *>(lane->d_scratch + 2 * kN);  // right behind the [2][N] dot-product buffer
    cuda_throw(launch_decrypt(lane->d_a, d_sk, lane->d_scratch, lane->d_plain, 1, s, d_flag), "decrypt");
    cuda_throw(cudaMemcpyAsync(lane->h_plain, lane->d_plain, kN * 2, cudaMemcpyDeviceToHost, s), "D2H plain");
    cuda_throw(cudaMemcpyAsync(lane->h_out, d_flag, sizeof(int), cudaMemcpyDeviceToHost, s), "D2H noise flag");
    cuda_throw(cudaStreamSynchronize(s), "stream sync");
    // sunscreen's Runtime::decrypt refuses a ciphertext whose invariant noise budget is 0 (-> fhe.rs:643, 696)
    if (*reinterpret_cast<const int *>(lane->h_out)) return kErrFailedDecryption;
    std::vector<uint8_t> scalar;
    decode_scalar(kind, lane->h_plain, kN, &scalar);
    // seed = SHA-512(public_data || whole input || plain bytes)   (fhe.rs:676, 646-649)
    std::vector<uint8_t> msg(public_data.p, public_data.p + public_data.n);
    msg.insert(msg.end(), in.p, in.p + in.n);
    msg.insert(msg.end(), scalar.begin(), scalar.end());
    return encrypt_plain(kind, Span{scalar.data(), scalar.size()}, pk, seed_from_hash(msg).w, &view, lane, out);
}

}  // namespace fheb
