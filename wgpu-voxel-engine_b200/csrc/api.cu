// api.cu -- the extern "C" boundary of libvoxb200.so (include/voxb200.h).
//
// Host side of the B200 path: one context owns one or more GPUs; each GPU has
// a compute stream, a copy stream, grow-only scratch buffers and two instance
// arenas.  Batch calls shard the coordinate list across the GPUs by contiguous
// index range with no inter-GPU traffic: a neighbour that lives in another
// shard (or outside the list but declared present) is a "halo" whose single
// border slab is regenerated from the noise on the GPU that needs it.
// Instances stream back through the two arenas (emit slice k+1 overlaps the
// D2H of slice k); into pageable caller memory they go through pinned bounce
// buffers and copy threads (ParallelCopier, d2h_pageable).  Lists of up to 32
// chunks -- the reference's one-chunk-per-call shape -- lease a scratch lane
// and take a one-synchronisation path (gen_mesh_short, mesh_short).
// There is no CPU compute path in this file: the host only moves bytes.
#include <algorithm>
#include <atomic>
#include <deque>
#include <functional>
#include <thread>
#include <unordered_map>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "host_copy.hpp"
#include "vx_common.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

int cuda_fail(cudaError_t e, const char *what, int line)
{
    cudaGetLastError();   // clear sticky-less errors
    return fail(e == cudaErrorMemoryAllocation ? VX_ERR_OOM : VX_ERR_CUDA, "CUDA error %d (%s) at api.cu:%d: %s",
                (int)e, cudaGetErrorString(e), line, what);
}

#define CK(expr)                                                      \
    do {                                                              \
        cudaError_t e_ = (expr);                                      \
        if (e_ != cudaSuccess) return cuda_fail(e_, #expr, __LINE__); \
    } while (0)
#define RET(expr)                 \
    do {                          \
        int r_ = (expr);          \
        if (r_ != VX_OK) return r_; \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);   // implicit device sync: nothing in flight still uses it
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

constexpr size_t kArenaBytes = size_t(256) << 20;                      // per instance arena
constexpr size_t kSliceBytes = size_t(256) << 20;                      // staging for fill / frustum outputs
constexpr size_t kMaxLanes = 8;                                        // concurrent short-list calls per context
constexpr size_t kShortList = 32;                                      // lists up to this long take the one-sync path
constexpr size_t kStageBytes = 4096;                                   // >= kShortList * (12 + 1 + 4 + 8) + 8 + 16
constexpr size_t kHostInstBytes = size_t(8) << 20;                     // pinned instance staging of the short-list path
// pinned bounce buffers (two) towards pageable memory, and the threads that copy out of them (default: half the
// host's hardware threads, at most 8).  Environment: VX_COPY_THREADS (documented in voxb200.h), VX_BOUNCE_KB (A/B aid).
// Measured on a 16-core host, one call of 1024 chunks = 319 MB into FRESH pageable memory: 2.1 GB/s with
// cudaMemcpyAsync straight into it, 2.2 / 7.4 / 12.1 / 15.6 GB/s through 4 MB bounce buffers with 1 / 4 / 8 / 16 threads;
// into already-touched pageable memory 16.7 -> 24-28 GB/s; into pinned memory 48-52 GB/s either way.
static const size_t kBounceBytes = [] {
    const char *e = getenv("VX_BOUNCE_KB");
    const long kb = e ? atol(e) : 4096;
    return (size_t)std::min(std::max(kb, 64L), 262144L) << 10;      // 64 KB .. 256 MB
}();
static const unsigned kCopyThreads = [] {
    const char *e = getenv("VX_COPY_THREADS");
    const long t = e ? atol(e) : 0;
    return (unsigned)std::min(std::max(t, 0L), 64L);                   // 0 = default
}();
// Which downloads into pageable memory take the bounce path (measured on a 16-core host, tools/probes/
// bounce_threshold_probe.py and pool_probe.cpp): cudaMemcpyAsync's own pageable path does 12-17 GB/s into pages that
// have been touched before, up to ~16 MB per copy, and collapses to 2-5 GB/s beyond that or when the pages are fresh
// (never faulted in); the bounce path does 10-16 GB/s into fresh pages and 24-28 GB/s into touched ones, but pays
// ~100 us of thread wake-ups per piece, which makes it the slower one below ~16 MB of touched memory.  Hence:
// everything from kBounceAlways up, and from kBounceMin up when the destination's pages are not resident yet.
// (VX_BOUNCE_MIN_KB / VX_BOUNCE_ALWAYS_KB: A/B aids.)
static const size_t kBounceMin = [] {
    const char *e = getenv("VX_BOUNCE_MIN_KB");
    const long kb = e ? atol(e) : 1024;
    return (size_t)std::max(kb, 64L) << 10;
}();
static const size_t kBounceAlways = [] {
    const char *e = getenv("VX_BOUNCE_ALWAYS_KB");
    const long kb = e ? atol(e) : 24576;
    return (size_t)std::max(kb, 64L) << 10;
}();
constexpr size_t kHostUpBytes = kShortList * (VX_VOXELS + 12 + 16 + 1) + 64;   // blocks, coords, nbr_idx, mesh_mask of a short list

struct Dev {
    int id = 0;
    int n_sm = 148;
    cudaStream_t stream = nullptr, copy = nullptr, aux = nullptr;
    cudaEvent_t emit_done[2] = {nullptr, nullptr}, d2h_done[2] = {nullptr, nullptr};
    cudaEvent_t fork = nullptr, join = nullptr;   // aux stream: neighbour resolve + halo fill beside the main fill
    DevBuf coords, mask8, occ, typ, halo, halo_items, table, nbr, nbr_y, nbr_in, counts, offsets, small, blocks, noise,
        inst[2], planes, vis, bits, mesh_mask;
    uint32_t *h_small = nullptr;   // pinned mirror of `small`
    unsigned char *h_stage = nullptr;   // pinned up/download staging of the short-list path (kStageBytes)
    unsigned char *h_inst = nullptr;    // pinned instance staging for pageable callers of that path (kHostInstBytes, on demand)
    unsigned char *h_up = nullptr;      // pinned upload staging of short vx_mesh lists (kHostUpBytes, on demand)
    unsigned char *h_bounce[2] = {nullptr, nullptr};   // pinned bounce buffers towards pageable destinations (on demand)
    cudaEvent_t bounce_ev[2] = {nullptr, nullptr};
    std::unique_ptr<ParallelCopier> copier;            // ... and the threads that copy out of them (on demand)
    bool is_lane = false;                              // a scratch lane: its caller thread copies alone
    vx_counter launches{0};
    size_t last_n = 0;             // chunks of the last device batch whose scratch is still valid (vx_emit_again_device)
    bool last_typ = false;         // ... and whether it had a DIRT plane
    bool ex_vertical = false;      // extension state of the current batch (vx_terrain_opts): nbr_y is valid
    int ex_dirt = 0;               // ... and the DIRT rule applies
    // optional per-kernel-class CUDA-event timing (vx_set_timing)
    std::atomic<bool> timing{false};       // read by lane callers without the context lock
    struct Stamp { cudaEvent_t a, b; int kind; };
    std::vector<Stamp> stamps;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> free_events;
    // `small` layout (bytes): 0 status u32 | 4 halo_count u32 | 8 near_zero u64 | 16 count-kernel ticket u32
    uint32_t *st() const { return small.as<uint32_t>(); }
    uint32_t *halo_count() const { return small.as<uint32_t>() + 1; }
    uint32_t *ticket() const { return small.as<uint32_t>() + 4; }
    unsigned long long *near() const { return reinterpret_cast<unsigned long long *>(small.as<uint8_t>() + 8); }
};

// records a CUDA-event pair around one kernel class on the launching stream
struct TimerScope {
    Dev &d;
    cudaStream_t s;
    Dev::Stamp st{nullptr, nullptr, 0};
    TimerScope(Dev &dev, cudaStream_t stream, int kind) : d(dev), s(stream)
    {
        if (!d.timing) return;
        if (d.free_events.empty()) {
            cudaEvent_t a, b;
            if (cudaEventCreate(&a) != cudaSuccess) return;
            if (cudaEventCreate(&b) != cudaSuccess) { cudaEventDestroy(a); return; }
            d.free_events.emplace_back(a, b);
        }
        st.a = d.free_events.back().first;
        st.b = d.free_events.back().second;
        st.kind = kind;
        d.free_events.pop_back();
        cudaEventRecord(st.a, s);
    }
    ~TimerScope()
    {
        if (!st.a) return;
        cudaEventRecord(st.b, s);
        d.stamps.push_back(st);
    }
};
#define TIMED(dev, stream, kind, call)       \
    do {                                     \
        TimerScope ts_(dev, stream, kind);   \
        CK(call);                            \
    } while (0)

}  // namespace

struct vx_ctx {
    std::deque<Dev> devs;          // (Dev holds an atomic: never copied or moved)
    std::mutex mu;
    // Lanes: extra scratch sets + streams on the GPU of devs[0], created on demand (at most kMaxLanes).  A
    // short-list vx_gen_mesh call leases one instead of taking `mu`, so the reference's pool threads
    // (chunk_builder.rs:80-84, one chunk per call on one shared handle) overlap on the GPU instead of queueing.
    std::vector<std::unique_ptr<Dev>> lanes;
    std::vector<char> lane_busy;
    std::mutex lane_mu;
    std::condition_variable lane_cv;
};

namespace {

int check_params(const vx_noise_params *in, vx_noise_params *out)
{
    if (in) *out = *in;
    else vx_default_noise_params(out);
    if (out->simplex_scale == 0.0f) out->simplex_scale = VX_SIMPLEX_SCALE_DEFAULT;
    if (out->octaves < 1 || out->octaves > VX_MAX_OCTAVES)
        return fail(VX_ERR_BAD_ARG, "octaves must be in 1..%d (got %d)", VX_MAX_OCTAVES, out->octaves);
    return VX_OK;
}

uint32_t table_size_for(size_t n)
{
    uint32_t t = 64;
    while ((size_t)t < 2 * n + 2) t <<= 1;
    return t;
}

// ---- device-side pipelines (everything async on stream s) -----------------

// resolve neighbours of chunks [lo, hi) of the list into d.nbr
int dev_resolve(Dev &d, cudaStream_t s, const int32_t *d_coords_all, uint32_t n_all, uint32_t lo, uint32_t hi,
                const uint8_t *d_mask_all, bool allow_halo, bool vertical = false)
{
    const uint32_t n = hi - lo;
    const size_t slots = vertical ? 6 : 4;          // neighbour slots (halo slabs at most) per chunk
    CK(d.nbr.ensure((size_t)n * 16));
    if (vertical) CK(d.nbr_y.ensure((size_t)n * 8));
    const uint32_t ts = table_size_for(n_all);
    CK(d.table.ensure((size_t)ts * 4));
    if (allow_halo) {
        CK(d.halo_items.ensure((size_t)n * slots * sizeof(vx_halo_item)));
        CK(d.halo.ensure((size_t)n * slots * 32 * 4));
    }
    vx_resolve_args r{};
    r.coords = d_coords_all;
    r.nbr_mask = d_mask_all;
    r.n = n_all;
    r.lo = lo;
    r.hi = hi;
    r.table = d.table.as<int32_t>();
    r.table_size = ts;
    r.nbr = d.nbr.as<int32_t>();
    r.nbr_y = vertical ? d.nbr_y.as<int32_t>() : nullptr;
    r.halo_items = allow_halo ? d.halo_items.as<vx_halo_item>() : nullptr;
    r.halo_count = d.halo_count();
    TIMED(d, s, VX_K_RESOLVE, vx_launch_resolve(r, s, &d.launches));
    return VX_OK;
}

// noise fill of chunks [lo, hi) + their halo slabs, then count + scan
int dev_gen_count(Dev &d, cudaStream_t s, const int32_t *d_coords_all, uint32_t n_all, uint32_t lo, uint32_t hi,
                  const vx_noise_params &p, const uint8_t *d_mask_all, uint32_t *d_counts, uint64_t *d_offsets,
                  uint8_t *d_blocks, const vx_terrain_opts *opts = nullptr)
{
    const uint32_t n = hi - lo;
    d.ex_vertical = opts && opts->vertical_neighbours;
    d.ex_dirt = (opts && opts->dirt_rule) ? 1 : 0;
    const bool ex = d.ex_vertical || d.ex_dirt;
    if (n == 0) return VX_OK;
    CK(d.occ.ensure((size_t)n * VX_ROWS * 4));
    const bool may_halo = d_mask_all != nullptr || lo != 0 || hi != n_all;
    // Neighbour resolution and the halo slabs depend only on the coordinates: they run on the
    // auxiliary stream beside the (400 us) main fill and join before the count kernel.
    CK(cudaEventRecord(d.fork, s));
    CK(cudaStreamWaitEvent(d.aux, d.fork, 0));
    RET(dev_resolve(d, d.aux, d_coords_all, n_all, lo, hi, d_mask_all, may_halo, d.ex_vertical));
    if (may_halo) {
        vx_fill_args h{};
        h.halo_items = d.halo_items.as<vx_halo_item>();
        h.halo_count = d.halo_count();
        h.max_units = n * (d.ex_vertical ? 6u : 4u);
        h.p = p;
        h.occ = d.halo.as<uint32_t>();
        TIMED(d, d.aux, VX_K_HALO, vx_launch_fill(h, true, d.aux, &d.launches));
    }
    CK(cudaEventRecord(d.join, d.aux));

    vx_fill_args f{};
    f.coords = d_coords_all + 3 * (size_t)lo;
    f.max_units = n * 32u;
    f.p = p;
    f.occ = d.occ.as<uint32_t>();
    f.blocks = ex ? nullptr : d_blocks;          // with the extension the types depend on the neighbours: written below
    TIMED(d, s, VX_K_FILL, vx_launch_fill(f, false, s, &d.launches));
    CK(cudaStreamWaitEvent(s, d.join, 0));
    vx_mesh_args m{};
    m.occ = d.occ.as<uint32_t>();
    m.typ = nullptr;
    m.halo = d.halo.as<uint32_t>();
    m.coords = d_coords_all + 3 * (size_t)lo;
    m.nbr = d.nbr.as<int32_t>();
    m.nbr_y = d.ex_vertical ? d.nbr_y.as<int32_t>() : nullptr;
    m.dirt_rule = d.ex_dirt;
    m.n = n;
    if (ex && d_blocks) TIMED(d, s, VX_K_PACK, vx_launch_blocks_ex(m, d_blocks, s, &d.launches));
    m.counts = d_counts;
    m.offsets = d_offsets;
    m.n = n;
    m.ticket = d.ticket();
    TIMED(d, s, VX_K_COUNT, vx_launch_mesh_count(m, s, &d.launches));     // the last CTA also scans
    return VX_OK;
}

vx_mesh_args emit_args(const Dev &d, const int32_t *d_coords, const uint32_t *typ, const uint8_t *mesh_mask,
                       uint32_t *d_counts, uint64_t *d_offsets, uint32_t n)
{
    vx_mesh_args m{};
    m.occ = d.occ.as<uint32_t>();
    m.typ = typ;
    m.halo = d.halo.as<uint32_t>();
    m.coords = d_coords;
    m.nbr = d.nbr.as<int32_t>();
    m.nbr_y = (!typ && d.ex_vertical) ? d.nbr_y.as<int32_t>() : nullptr;
    m.dirt_rule = typ ? 0 : d.ex_dirt;
    m.mesh_mask = mesh_mask;
    m.counts = d_counts;
    m.offsets = d_offsets;
    m.n = n;
    return m;
}

// true for memory cudaMemcpyAsync can reach without the driver's pageable staging (pinned / registered / managed)
bool host_is_pinned(const void *p)
{
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type != cudaMemoryTypeUnregistered;
}

// does a download of `bytes` into `dst` take the bounce path?
bool use_bounce(const void *dst, size_t bytes)
{
    if (bytes < kBounceMin || host_is_pinned(dst)) return false;
    return bytes >= kBounceAlways || host_is_fresh(dst, bytes);
}

// Device -> PAGEABLE host memory: pieces of kBounceBytes go through two pinned bounce buffers on stream s while the
// copier threads move the previous piece to its destination.  Returns when all bytes are in dst.
int d2h_pageable(Dev &d, cudaStream_t s, void *dst_v, const void *src_v, size_t bytes)
{
    unsigned char *dst = static_cast<unsigned char *>(dst_v);
    const unsigned char *src = static_cast<const unsigned char *>(src_v);
    for (int b = 0; b < 2; ++b) {
        if (!d.h_bounce[b]) CK(cudaHostAlloc((void **)&d.h_bounce[b], kBounceBytes, cudaHostAllocPortable));
        if (!d.bounce_ev[b]) CK(cudaEventCreateWithFlags(&d.bounce_ev[b], cudaEventDisableTiming));
    }
    if (!d.copier) {
        // lanes serve pool threads that already run side by side: no helpers there
        const unsigned hw = std::max(2u, std::thread::hardware_concurrency());
        d.copier.reset(new ParallelCopier(d.is_lane ? 0u : (kCopyThreads ? kCopyThreads : std::min(8u, hw / 2)) - 1));
    }
    size_t issued = 0, copied = 0;
    int k = 0;
    while (copied < bytes) {
        if (issued < bytes) {
            const size_t len = std::min(kBounceBytes, bytes - issued);
            CK(cudaMemcpyAsync(d.h_bounce[k & 1], src + issued, len, cudaMemcpyDeviceToHost, s));
            CK(cudaEventRecord(d.bounce_ev[k & 1], s));
            issued += len;
            ++k;
        }
        // the piece before the one just issued (or the last one) is complete or about to be: move it out
        if (issued - copied > kBounceBytes || issued == bytes) {
            const int b = (int)((copied / kBounceBytes) & 1);
            const size_t len = std::min(kBounceBytes, bytes - copied);
            CK(cudaEventSynchronize(d.bounce_ev[b]));
            d.copier->copy(dst + copied, d.h_bounce[b], len);
            copied += len;
        }
    }
    return VX_OK;
}

// device -> host on stream s: asynchronous into pinned (or small) destinations, through the bounce buffers -- and then
// complete on return -- into large pageable ones
int d2h_any(Dev &d, cudaStream_t s, void *dst, const void *src, size_t bytes)
{
    if (use_bounce(dst, bytes)) return d2h_pageable(d, s, dst, src, bytes);
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
    return VX_OK;
}

// Stream the instances of chunks [0, n) (device-local numbering, host copy of
// the local offsets in h_off) to host memory through the two arenas.
int dev_stream_instances(Dev &d, vx_mesh_args m, const uint64_t *h_off, uint32_t n, void *h_out_v,
                         bool packed = false)
{
    const size_t rec = packed ? sizeof(uint32_t) : sizeof(vx_instance);
    const size_t kArenaQuads = kArenaBytes / rec;
    unsigned char *h_out = static_cast<unsigned char *>(h_out_v);
    m.packed = packed ? 1 : 0;
    // a list whose instances fit one arena only needs that much of it (lanes serve one-chunk calls)
    const size_t need = std::min(kArenaBytes, std::max<size_t>((size_t)h_off[n] * rec, 256));
    CK(d.inst[0].ensure(need));
    if (h_off[n] > kArenaQuads) CK(d.inst[1].ensure(kArenaBytes));
    // pageable destination: every slice is drained through the pinned bounce buffers (d2h_pageable) AFTER the next
    // slice's emit has been launched, so the GPU keeps working while the host copies
    const bool pageable = use_bounce(h_out, h_off[n] * rec);
    struct Pending { int buf; uint64_t first, quads; bool valid; } prev{0, 0, 0, false};
    auto drain = [&](const Pending &pd) -> int {
        CK(cudaStreamWaitEvent(d.copy, d.emit_done[pd.buf], 0));
        RET(d2h_pageable(d, d.copy, h_out + pd.first * rec, d.inst[pd.buf].p, pd.quads * rec));
        CK(cudaEventRecord(d.d2h_done[pd.buf], d.copy));
        return VX_OK;
    };
    uint32_t a = 0;
    int k = 0;
    while (a < n) {
        uint32_t b = a + 1;     // at least one chunk (<= 98304 quads always fits)
        while (b < n && h_off[b + 1] - h_off[a] <= kArenaQuads) ++b;
        const uint64_t q = h_off[b] - h_off[a];
        if (q > 0) {
            const int buf = k & 1;
            CK(cudaStreamWaitEvent(d.stream, d.d2h_done[buf], 0));      // arena free again
            m.out = d.inst[buf].as<vx_instance>();
            m.capacity = std::min<size_t>(kArenaQuads, d.inst[buf].cap / rec);
            m.m0 = a;
            m.quad_base = h_off[a];
            m.status = d.st();
            TIMED(d, d.stream, VX_K_EMIT, vx_launch_mesh_emit(m, b - a, d.stream, &d.launches));
            CK(cudaEventRecord(d.emit_done[buf], d.stream));
            if (pageable) {
                if (prev.valid) RET(drain(prev));
                prev = Pending{buf, h_off[a], q, true};
            } else {
                CK(cudaStreamWaitEvent(d.copy, d.emit_done[buf], 0));
                CK(cudaMemcpyAsync(h_out + h_off[a] * rec, d.inst[buf].p, q * rec, cudaMemcpyDeviceToHost, d.copy));
                CK(cudaEventRecord(d.d2h_done[buf], d.copy));
            }
            ++k;
        }
        a = b;
    }
    if (prev.valid) RET(drain(prev));
    return VX_OK;
}

int sync_dev(Dev &d)
{
    CK(cudaSetDevice(d.id));
    CK(cudaStreamSynchronize(d.stream));
    CK(cudaStreamSynchronize(d.copy));
    return VX_OK;
}

void split_range(size_t n, size_t nd, size_t d, size_t align, uint32_t *lo, uint32_t *hi)
{
    size_t per = (n + nd - 1) / nd;
    per = (per + align - 1) / align * align;
    size_t l = std::min(n, per * d), h = std::min(n, per * (d + 1));
    *lo = (uint32_t)l;
    *hi = (uint32_t)h;
}

cudaError_t init_dev(Dev &d, int id)
{
    d.id = id;
    cudaError_t e = cudaSetDevice(d.id);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&d.n_sm, cudaDevAttrMultiProcessorCount, d.id);
    if (e == cudaSuccess) e = vx_init_mesh_kernels();
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.copy, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.aux, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.join, cudaEventDisableTiming);
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
        e = cudaEventCreateWithFlags(&d.emit_done[b], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.d2h_done[b], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = d.small.ensure(64);
    if (e == cudaSuccess) e = cudaMemset(d.small.p, 0, 64);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&d.h_small, 64, cudaHostAllocPortable);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&d.h_stage, kStageBytes, cudaHostAllocPortable);
    return e;
}

void destroy_dev(Dev &d)
{
    cudaSetDevice(d.id);
    cudaDeviceSynchronize();
    for (DevBuf *b : {&d.coords, &d.mask8, &d.occ, &d.typ, &d.halo, &d.halo_items, &d.table, &d.nbr, &d.nbr_y, &d.nbr_in,
                      &d.counts, &d.offsets, &d.small, &d.blocks, &d.noise, &d.inst[0], &d.inst[1], &d.planes,
                      &d.vis, &d.bits, &d.mesh_mask})
        b->release();
    if (d.h_small) cudaFreeHost(d.h_small);
    if (d.h_stage) cudaFreeHost(d.h_stage);
    if (d.h_inst) cudaFreeHost(d.h_inst);
    if (d.h_up) cudaFreeHost(d.h_up);
    d.copier.reset();
    for (int b = 0; b < 2; ++b) {
        if (d.h_bounce[b]) cudaFreeHost(d.h_bounce[b]);
        if (d.bounce_ev[b]) cudaEventDestroy(d.bounce_ev[b]);
    }
    for (auto &st : d.stamps) d.free_events.emplace_back(st.a, st.b);
    for (auto &ev : d.free_events) {
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    for (int b = 0; b < 2; ++b) {
        if (d.emit_done[b]) cudaEventDestroy(d.emit_done[b]);
        if (d.d2h_done[b]) cudaEventDestroy(d.d2h_done[b]);
    }
    if (d.stream) cudaStreamDestroy(d.stream);
    if (d.copy) cudaStreamDestroy(d.copy);
    if (d.aux) cudaStreamDestroy(d.aux);
    if (d.fork) cudaEventDestroy(d.fork);
    if (d.join) cudaEventDestroy(d.join);
}

// Runs fn(device index) for every device of a multi-GPU context, one host thread per device, so that a
// blocking step on one GPU (a D2H into pageable caller memory only returns when the copy is done) does
// not hold back the launches of the others.  Errors keep their text: vx_last_error is thread-local.
int for_each_device(size_t nd, const std::function<int(size_t)> &fn)
{
    if (nd == 1) return fn(0);
    std::vector<int> rc(nd, VX_OK);
    std::vector<std::string> msg(nd);
    std::vector<std::thread> th;
    for (size_t di = 0; di < nd; ++di)
        th.emplace_back([&, di] {
            rc[di] = fn(di);
            if (rc[di] != VX_OK) msg[di] = g_err;
        });
    for (auto &t : th) t.join();
    for (size_t di = 0; di < nd; ++di)
        if (rc[di] != VX_OK) {
            g_err = msg[di];
            return rc[di];
        }
    return VX_OK;
}

// Lease of one lane for the duration of a short-list call.
struct LaneLease {
    vx_ctx *ctx = nullptr;
    int idx = -1;
    Dev *d = nullptr;
    cudaError_t err = cudaSuccess;
    explicit LaneLease(vx_ctx *c) : ctx(c)
    {
        std::unique_lock<std::mutex> lk(ctx->lane_mu);
        for (;;) {
            for (size_t i = 0; i < ctx->lanes.size(); ++i)
                if (!ctx->lane_busy[i]) {
                    idx = (int)i;
                    break;
                }
            if (idx >= 0) break;
            if (ctx->lanes.size() < kMaxLanes) {
                std::unique_ptr<Dev> nd(new Dev());
                nd->is_lane = true;
                err = init_dev(*nd, ctx->devs[0].id);
                if (err != cudaSuccess) {
                    destroy_dev(*nd);
                    return;
                }
                ctx->lanes.push_back(std::move(nd));
                ctx->lane_busy.push_back(0);
                idx = (int)ctx->lanes.size() - 1;
                break;
            }
            ctx->lane_cv.wait(lk);
        }
        ctx->lane_busy[idx] = 1;
        d = ctx->lanes[idx].get();
    }
    ~LaneLease()
    {
        if (idx < 0) return;
        {
            std::lock_guard<std::mutex> lk(ctx->lane_mu);
            ctx->lane_busy[idx] = 0;
        }
        ctx->lane_cv.notify_one();
    }
};

// The Dev a host call runs on: a leased lane for short lists (concurrent callers overlap), else the
// context's first device under the context lock.
struct DevGuard {
    std::unique_lock<std::mutex> lk;
    std::unique_ptr<LaneLease> lease;
    Dev *d = nullptr;
    cudaError_t err = cudaSuccess;
    DevGuard(vx_ctx *ctx, size_t n)
    {
        if (n <= kShortList && !ctx->devs[0].timing) {
            lease.reset(new LaneLease(ctx));
            d = lease->d;
            err = lease->err;
        } else {
            lk = std::unique_lock<std::mutex>(ctx->mu);
            d = &ctx->devs[0];
        }
    }
};

}  // namespace

// ===========================================================================
extern "C" {

int vx_version(void) { return VX_VERSION; }

const char *vx_last_error(void) { return g_err.c_str(); }

int vx_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(VX_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
    }
    return n;
}

void vx_default_noise_params(vx_noise_params *p)
{
    // perlin_noise.rs:14-18
    p->seed = 1;
    p->freq = 0.09f;
    p->gain = 2.5f;
    p->lacunarity = 0.5f;
    p->octaves = 3;
    p->simplex_scale = VX_SIMPLEX_SCALE_DEFAULT;
}

int vx_create(const int *devices, int n_devices, vx_ctx **out)
{
    if (!out) return fail(VX_ERR_BAD_ARG, "out is NULL");
    *out = nullptr;
    const int avail = vx_device_count();
    if (avail <= 0) return fail(VX_ERR_NO_DEVICE, "no CUDA device available (libvoxb200 has no CPU fallback)");
    std::vector<int> ids;
    if (devices) {
        if (n_devices <= 0) return fail(VX_ERR_BAD_ARG, "n_devices must be > 0 when devices is given");
        ids.assign(devices, devices + n_devices);
    } else {
        const int n = n_devices <= 0 ? avail : n_devices;
        for (int i = 0; i < n; ++i) ids.push_back(i);
    }
    for (int id : ids)
        if (id < 0 || id >= avail) return fail(VX_ERR_BAD_ARG, "device %d out of range (have %d)", id, avail);
    vx_ctx *ctx = new vx_ctx();
    for (size_t i = 0; i < ids.size(); ++i) ctx->devs.emplace_back();
    for (size_t i = 0; i < ids.size(); ++i) {
        const cudaError_t e = init_dev(ctx->devs[i], ids[i]);
        if (e != cudaSuccess) {
            int r = cuda_fail(e, "context setup", __LINE__);
            vx_destroy(ctx);
            return r;
        }
    }
    *out = ctx;
    return VX_OK;
}

void vx_destroy(vx_ctx *ctx)
{
    if (!ctx) return;
    for (Dev &d : ctx->devs) destroy_dev(d);
    for (auto &l : ctx->lanes) destroy_dev(*l);
    delete ctx;
}

int vx_num_devices(const vx_ctx *ctx) { return ctx ? (int)ctx->devs.size() : 0; }

uint64_t vx_launch_count(const vx_ctx *ctx)
{
    uint64_t t = 0;
    if (ctx) {
        for (const Dev &d : ctx->devs) t += d.launches.load(std::memory_order_relaxed);
        std::lock_guard<std::mutex> lk(const_cast<vx_ctx *>(ctx)->lane_mu);
        for (const auto &l : ctx->lanes) t += l->launches.load(std::memory_order_relaxed);
    }
    return t;
}

int vx_host_alloc(void **out, size_t bytes)
{
    if (!out) return fail(VX_ERR_BAD_ARG, "out is NULL");
    if (vx_device_count() <= 0) return VX_ERR_NO_DEVICE;
    CK(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return VX_OK;
}

int vx_host_free(void *p)
{
    if (p) CK(cudaFreeHost(p));
    return VX_OK;
}

int vx_sync(vx_ctx *ctx)
{
    if (!ctx) return fail(VX_ERR_BAD_ARG, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (Dev &d : ctx->devs) RET(sync_dev(d));
    return VX_OK;
}

// ---------------------------------------------------------------------------
// fill
// ---------------------------------------------------------------------------
int vx_fill_device(vx_ctx *ctx, const int32_t *coords, size_t n, const vx_noise_params *params,
                   uint8_t *blocks_out, float *noise_out, uint32_t *occupancy_out, uint64_t *near_zero_out,
                   void *stream)
{
    if (!ctx) return fail(VX_ERR_BAD_ARG, "ctx is NULL");
    if (n == 0) return VX_OK;
    if (!coords) return fail(VX_ERR_BAD_ARG, "coords is NULL");
    if (n > (size_t)0x7ffffff) return fail(VX_ERR_BAD_ARG, "n too large for one call (%zu)", n);
    vx_noise_params p;
    RET(check_params(params, &p));
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dev &d = ctx->devs[0];
    CK(cudaSetDevice(d.id));
    cudaStream_t s = stream ? (cudaStream_t)stream : d.stream;
    uint32_t *occ = occupancy_out;
    if (!occ) {
        CK(d.occ.ensure(n * VX_ROWS * 4));
        occ = d.occ.as<uint32_t>();
    }
    vx_fill_args f{};
    f.coords = coords;
    f.max_units = (uint32_t)n * 32u;
    f.p = p;
    f.occ = occ;
    f.blocks = blocks_out;
    f.noise = noise_out;
    f.near_zero = reinterpret_cast<unsigned long long *>(near_zero_out);
    if (near_zero_out) CK(cudaMemsetAsync(near_zero_out, 0, 8, s));
    TIMED(d, s, VX_K_FILL, vx_launch_fill(f, false, s, &d.launches));
    return VX_OK;
}

int vx_fill(vx_ctx *ctx, const int32_t *coords, size_t n, const vx_noise_params *params, uint8_t *blocks_out,
            float *noise_out, uint32_t *occupancy_out, uint64_t *near_zero_out)
{
    if (!ctx) return fail(VX_ERR_BAD_ARG, "ctx is NULL");
    if (near_zero_out) *near_zero_out = 0;
    if (n == 0) return VX_OK;
    if (!coords) return fail(VX_ERR_BAD_ARG, "coords is NULL");
    if (n > (size_t)0x7ffffff) return fail(VX_ERR_BAD_ARG, "n too large for one call (%zu)", n);
    vx_noise_params p;
    RET(check_params(params, &p));
    DevGuard guard(ctx, n);
    if (!guard.d) return cuda_fail(guard.err, "lane setup", __LINE__);
    const bool on_lane = guard.lease != nullptr;
    const size_t nd = on_lane ? 1 : ctx->devs.size();
    const size_t per_chunk = VX_ROWS * 4 + (blocks_out ? VX_VOXELS : 0) + (noise_out ? VX_VOXELS * 4 : 0);
    const size_t slice = std::max<size_t>(1, kSliceBytes / per_chunk);
    for (size_t di = 0; di < nd; ++di) {
        Dev &d = on_lane ? *guard.d : ctx->devs[di];
        uint32_t lo, hi;
        split_range(n, nd, di, 1, &lo, &hi);
        if (hi <= lo) continue;
        CK(cudaSetDevice(d.id));
        const size_t nn = hi - lo, sl = std::min(slice, nn);
        CK(d.coords.ensure(nn * 12));
        CK(cudaMemcpyAsync(d.coords.p, coords + 3 * (size_t)lo, nn * 12, cudaMemcpyHostToDevice, d.stream));
        CK(d.occ.ensure(sl * VX_ROWS * 4));
        if (blocks_out) CK(d.blocks.ensure(sl * VX_VOXELS));
        if (noise_out) CK(d.noise.ensure(sl * VX_VOXELS * 4));
        if (near_zero_out) CK(cudaMemsetAsync(d.near(), 0, 8, d.stream));
        for (size_t a = 0; a < nn; a += sl) {
            const size_t c = std::min(sl, nn - a);
            vx_fill_args f{};
            f.coords = d.coords.as<int32_t>() + 3 * a;
            f.max_units = (uint32_t)c * 32u;
            f.p = p;
            f.occ = d.occ.as<uint32_t>();
            f.blocks = blocks_out ? d.blocks.as<uint8_t>() : nullptr;
            f.noise = noise_out ? d.noise.as<float>() : nullptr;
            f.near_zero = near_zero_out ? d.near() : nullptr;
            TIMED(d, d.stream, VX_K_FILL, vx_launch_fill(f, false, d.stream, &d.launches));
            const size_t g = lo + a;
            if (blocks_out) RET(d2h_any(d, d.stream, blocks_out + g * VX_VOXELS, d.blocks.p, c * VX_VOXELS));
            if (noise_out) RET(d2h_any(d, d.stream, noise_out + g * VX_VOXELS, d.noise.p, c * VX_VOXELS * 4));
            if (occupancy_out) RET(d2h_any(d, d.stream, occupancy_out + g * VX_ROWS, d.occ.p, c * VX_ROWS * 4));
        }
        if (near_zero_out) CK(cudaMemcpyAsync(d.h_small + 2, d.near(), 8, cudaMemcpyDeviceToHost, d.stream));
    }
    for (size_t di = 0; di < nd; ++di) {
        Dev &d = on_lane ? *guard.d : ctx->devs[di];
        RET(sync_dev(d));
        if (near_zero_out) {
            uint64_t v;
            memcpy(&v, d.h_small + 2, 8);
            *near_zero_out += v;
            memset(d.h_small + 2, 0, 8);
        }
    }
    return VX_OK;
}

// ---------------------------------------------------------------------------
// mesh of caller-provided blocks (first GPU of the context)
// ---------------------------------------------------------------------------
static int mesh_prepare(Dev &d, cudaStream_t s, const int32_t *d_coords, uint32_t n, const uint8_t *d_blocks,
                        const int32_t *d_nbr_idx, uint32_t *d_counts, uint64_t *d_offsets,
                        const uint8_t *d_mesh_mask, uint32_t *d_status)
{
    d.ex_vertical = false;
    d.ex_dirt = 0;
    CK(d.occ.ensure((size_t)n * VX_ROWS * 4));
    CK(d.typ.ensure((size_t)n * VX_ROWS * 4));
    TIMED(d, s, VX_K_PACK, vx_launch_pack(d_blocks, n, d.occ.as<uint32_t>(), d.typ.as<uint32_t>(), d_status, s, &d.launches));
    if (d_nbr_idx) {
        CK(d.nbr.ensure((size_t)n * 16));
        TIMED(d, s, VX_K_RESOLVE, vx_launch_remap_nbr(d_nbr_idx, d.nbr.as<int32_t>(), n, d_status, s, &d.launches));
    } else {
        RET(dev_resolve(d, s, d_coords, n, 0, n, nullptr, false));
    }
    vx_mesh_args m = emit_args(d, d_coords, d.typ.as<uint32_t>(), d_mesh_mask, d_counts, d_offsets, n);
    m.ticket = d.ticket();
    TIMED(d, s, VX_K_COUNT, vx_launch_mesh_count(m, s, &d.launches));     // the last CTA also scans
    return VX_OK;
}

int vx_mesh_device(vx_ctx *ctx, const int32_t *coords, size_t n, const uint8_t *blocks, const int32_t *nbr_idx,
                   const uint8_t *mesh_mask, uint32_t *counts, uint64_t *offsets, vx_instance *instances,
                   size_t capacity, uint32_t *status, void *stream)
{
    if (!ctx) return fail(VX_ERR_BAD_ARG, "ctx is NULL");
    if (!offsets || (n && (!coords || !blocks || !counts))) return fail(VX_ERR_BAD_ARG, "NULL argument");
    if (n > (size_t)0x3ffffff) return fail(VX_ERR_BAD_ARG, "n too large for one call (%zu)", n);
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dev &d = ctx->devs[0];
    CK(cudaSetDevice(d.id));
    cudaStream_t s = stream ? (cudaStream_t)stream : d.stream;
    if (status) CK(cudaMemsetAsync(status, 0, 4, s));
    if (n == 0) {
        CK(cudaMemsetAsync(offsets, 0, 8, s));
        return VX_OK;
    }
    d.last_n = 0;
    RET(mesh_prepare(d, s, coords, (uint32_t)n, blocks, nbr_idx, counts, offsets, mesh_mask, status));
    d.last_n = mesh_mask ? 0 : n;          // (a masked batch is not replayable: the mask is the caller's)
    d.last_typ = true;
    if (instances) {
        vx_mesh_args m = emit_args(d, coords, d.typ.as<uint32_t>(), mesh_mask, counts, offsets, (uint32_t)n);
        m.out = instances;
        m.capacity = capacity;
        m.status = status;
        TIMED(d, s, VX_K_EMIT, vx_launch_mesh_emit(m, (uint32_t)n, s, &d.launches));
    }
    return VX_OK;
}

// vx_mesh over several GPUs: the list is split by contiguous index range; every device also receives, as
// mesh-less "ghost" chunks, the few chunks outside its range that its own chunks name as neighbours (for a
// contiguous band of a grid: one row on either side), so the device-side mesher needs nothing from its peers.
static int mesh_multi(vx_ctx *ctx, const int32_t *coords, size_t n, const uint8_t *blocks, const int32_t *nbr_idx,
                      const uint8_t *mesh_mask, uint32_t *counts, uint64_t *offsets, vx_instance *instances,
                      size_t capacity, uint64_t *total_out)
{
    const size_t nd = ctx->devs.size();
    std::lock_guard<std::mutex> lk(ctx->mu);
    // neighbour table of the whole list: the caller's, or looked up by coordinate (chunk_builder.rs:134-169)
    std::vector<int32_t> nbr_all;
    if (!nbr_idx) {
        struct Key { int32_t x, y, z; bool operator==(const Key &o) const { return x == o.x && y == o.y && z == o.z; } };
        struct KeyHash { size_t operator()(const Key &k) const {
            return ((uint64_t)(uint32_t)k.x * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)(uint32_t)k.z * 0xC2B2AE3D27D4EB4Full) ^
                   ((uint64_t)(uint32_t)k.y * 0xD6E8FEB86659FD93ull); } };
        std::unordered_map<Key, int32_t, KeyHash> where;
        where.reserve(n * 2);
        for (size_t i = 0; i < n; ++i) where.emplace(Key{coords[3 * i], coords[3 * i + 1], coords[3 * i + 2]}, (int32_t)i);   // first wins
        nbr_all.resize(n * 4);
        static const int dx[4] = {-1, 1, 0, 0}, dz[4] = {0, 0, -1, 1};
        for (size_t i = 0; i < n; ++i)
            for (int k = 0; k < 4; ++k) {
                const Key q{(int32_t)((uint32_t)coords[3 * i] + (uint32_t)dx[k]), coords[3 * i + 1],
                            (int32_t)((uint32_t)coords[3 * i + 2] + (uint32_t)dz[k])};
                auto it = where.find(q);
                nbr_all[4 * i + k] = it == where.end() ? -1 : it->second;
            }
        nbr_idx = nbr_all.data();
    }
    std::vector<uint32_t> lo(nd), hi(nd);
    std::vector<std::vector<uint64_t>> loc(nd);
    std::vector<int> bad_block(nd, 0);
    RET(for_each_device(nd, [&](size_t di) -> int {
        Dev &d = ctx->devs[di];
        split_range(n, nd, di, 1, &lo[di], &hi[di]);
        const uint32_t c = hi[di] - lo[di];
        if (c == 0) return VX_OK;
        // local list = own chunks, then ghosts
        std::vector<int32_t> ghosts;
        std::unordered_map<int32_t, int32_t> ghost_slot;
        std::vector<int32_t> l_nbr((size_t)c * 4);
        for (uint32_t i = 0; i < c; ++i)
            for (int k = 0; k < 4; ++k) {
                const int32_t g = nbr_idx[4 * (size_t)(lo[di] + i) + k];
                int32_t l = -1;
                if (g >= 0) {
                    if ((uint32_t)g >= lo[di] && (uint32_t)g < hi[di]) l = g - (int32_t)lo[di];
                    else {
                        auto it = ghost_slot.find(g);
                        if (it == ghost_slot.end()) {
                            it = ghost_slot.emplace(g, (int32_t)(c + ghosts.size())).first;
                            ghosts.push_back(g);
                        }
                        l = it->second;
                    }
                }
                l_nbr[4 * (size_t)i + k] = l;
            }
        const size_t m = (size_t)c + ghosts.size();
        l_nbr.resize(m * 4, -1);
        std::vector<int32_t> l_coords(m * 3);
        std::memcpy(l_coords.data(), coords + 3 * (size_t)lo[di], (size_t)c * 12);
        std::vector<uint8_t> l_mask(m, 0);
        for (uint32_t i = 0; i < c; ++i) l_mask[i] = mesh_mask ? mesh_mask[lo[di] + i] : 1;
        for (size_t j = 0; j < ghosts.size(); ++j) std::memcpy(&l_coords[3 * (c + j)], coords + 3 * (size_t)ghosts[j], 12);
        CK(cudaSetDevice(d.id));
        cudaStream_t s = d.stream;
        CK(d.coords.ensure(m * 12));
        CK(d.blocks.ensure(m * VX_VOXELS));
        CK(d.counts.ensure(m * 4));
        CK(d.offsets.ensure((m + 1) * 8));
        CK(d.nbr_in.ensure(m * 16));
        CK(d.mesh_mask.ensure(m));
        CK(cudaMemcpyAsync(d.coords.p, l_coords.data(), m * 12, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(d.blocks.p, blocks + (size_t)lo[di] * VX_VOXELS, (size_t)c * VX_VOXELS, cudaMemcpyHostToDevice, s));
        for (size_t j = 0; j < ghosts.size(); ++j)
            CK(cudaMemcpyAsync(d.blocks.as<uint8_t>() + (c + j) * VX_VOXELS, blocks + (size_t)ghosts[j] * VX_VOXELS,
                               VX_VOXELS, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(d.nbr_in.p, l_nbr.data(), m * 16, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(d.mesh_mask.p, l_mask.data(), m, cudaMemcpyHostToDevice, s));
        CK(cudaMemsetAsync(d.st(), 0, 4, s));
        RET(mesh_prepare(d, s, d.coords.as<int32_t>(), (uint32_t)m, d.blocks.as<uint8_t>(), d.nbr_in.as<int32_t>(),
                         d.counts.as<uint32_t>(), d.offsets.as<uint64_t>(), d.mesh_mask.as<uint8_t>(), d.st()));
        loc[di].resize((size_t)c + 1);
        CK(cudaMemcpyAsync(counts + lo[di], d.counts.p, (size_t)c * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(loc[di].data(), d.offsets.p, ((size_t)c + 1) * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(d.h_small, d.st(), 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));               // (the host vectors above must outlive the copies)
        bad_block[di] = (d.h_small[0] & VX_ST_BAD_BLOCK) ? 1 : 0;
        return VX_OK;
    }));
    for (size_t di = 0; di < nd; ++di)
        if (bad_block[di])
            return fail(VX_ERR_BAD_BLOCK, "blocks contain an id outside {0 GRASS, 1 DIRT, 2 AIR} (cube.rs:16-21)");
    uint64_t total = 0;
    std::vector<uint64_t> base(nd, 0);
    for (size_t di = 0; di < nd; ++di) {
        const uint32_t c = hi[di] - lo[di];
        if (c == 0) continue;
        base[di] = total;
        for (uint32_t i = 0; i < c; ++i) offsets[lo[di] + i] = total + loc[di][i];
        total += loc[di][c];
    }
    offsets[n] = total;
    if (total_out) *total_out = total;
    if (!instances) return VX_OK;
    if (total > capacity)
        return fail(VX_ERR_CAPACITY, "instance capacity %zu < %llu quads", capacity, (unsigned long long)total);
    return for_each_device(nd, [&](size_t di) -> int {
        const uint32_t c = hi[di] - lo[di];
        if (c == 0) return VX_OK;
        Dev &d = ctx->devs[di];
        CK(cudaSetDevice(d.id));
        // ghosts sit behind the own chunks and have count 0: emitting the first c chunks is the whole shard
        vx_mesh_args m = emit_args(d, d.coords.as<int32_t>(), d.typ.as<uint32_t>(), d.mesh_mask.as<uint8_t>(),
                                   d.counts.as<uint32_t>(), d.offsets.as<uint64_t>(), c);
        RET(dev_stream_instances(d, m, loc[di].data(), c, reinterpret_cast<unsigned char *>(instances) + base[di] * sizeof(vx_instance)));
        return sync_dev(d);
    });
}

// Short lists (Chunk::create_mesh of the reference: the chunk and its four neighbours, chunk.rs:48-57): the same
// one-synchronisation shape as gen_mesh_short.  Blocks, coordinates, neighbour wiring and mask go up in ONE copy from
// pinned staging (a copy from pageable caller memory would go through the driver's own staging, one caller at a
// time); the emit is launched speculatively behind the count into the caller's pinned buffer or the lane's pinned
// staging; offsets, counts and the status word come back in one copy.
static int mesh_short(Dev &d, const int32_t *coords, size_t n, const uint8_t *blocks, const int32_t *nbr_idx,
                      const uint8_t *mesh_mask, uint32_t *counts, uint64_t *offsets, unsigned char *instances,
                      size_t capacity, uint64_t *total_out)
{
    const size_t rec = sizeof(vx_instance);
    const uint32_t nn = (uint32_t)n;
    CK(cudaSetDevice(d.id));
    cudaStream_t s = d.stream;
    if (!d.h_up) CK(cudaHostAlloc((void **)&d.h_up, kHostUpBytes, cudaHostAllocPortable));
    // upload block: [blocks 32768 n][coords 12 n][nbr_idx 16 n][mesh_mask n]
    const size_t o_coords = n * VX_VOXELS, o_nbr = o_coords + 12 * n, o_mask = o_nbr + (nbr_idx ? 16 * n : 0);
    const size_t up = o_mask + (mesh_mask ? n : 0);
    CK(d.blocks.ensure(up));
    std::memcpy(d.h_up, blocks, n * VX_VOXELS);
    std::memcpy(d.h_up + o_coords, coords, 12 * n);
    if (nbr_idx) std::memcpy(d.h_up + o_nbr, nbr_idx, 16 * n);
    if (mesh_mask) std::memcpy(d.h_up + o_mask, mesh_mask, n);
    CK(cudaMemcpyAsync(d.blocks.p, d.h_up, up, cudaMemcpyHostToDevice, s));
    const unsigned char *base = d.blocks.as<unsigned char>();
    const int32_t *d_coords = reinterpret_cast<const int32_t *>(base + o_coords);
    const int32_t *d_nbr_idx = nbr_idx ? reinterpret_cast<const int32_t *>(base + o_nbr) : nullptr;
    const uint8_t *d_mm = mesh_mask ? base + o_mask : nullptr;
    // download block on the device: [offsets 8 (n + 1)][counts 4 n]
    const size_t off_bytes = 8 * (n + 1), down = off_bytes + 4 * n;
    CK(d.offsets.ensure(down));
    uint64_t *d_off = d.offsets.as<uint64_t>();
    uint32_t *d_cnt = reinterpret_cast<uint32_t *>(d.offsets.as<unsigned char>() + off_bytes);
    CK(cudaMemsetAsync(d.st(), 0, 4, s));
    RET(mesh_prepare(d, s, d_coords, nn, base, d_nbr_idx, d_cnt, d_off, d_mm, d.st()));
    void *dev_out = nullptr;
    size_t dev_cap = 0;
    bool staged = false;
    if (instances) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, instances) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
            dev_out = at.devicePointer;               // pinned caller buffer: the emit writes it directly
            dev_cap = capacity;
        } else {
            cudaGetLastError();                       // a pageable pointer is not an error here
            if (!d.h_inst && cudaHostAlloc((void **)&d.h_inst, kHostInstBytes, cudaHostAllocPortable) != cudaSuccess) {
                cudaGetLastError();
                d.h_inst = nullptr;
            }
            void *dp = nullptr;
            if (d.h_inst && cudaHostGetDevicePointer(&dp, d.h_inst, 0) == cudaSuccess) {
                dev_out = dp;
                dev_cap = std::min(capacity, kHostInstBytes / rec);
                staged = true;
            } else {
                cudaGetLastError();
            }
        }
        if (dev_out) {
            vx_mesh_args m = emit_args(d, d_coords, d.typ.as<uint32_t>(), d_mm, d_cnt, d_off, nn);
            m.out = static_cast<vx_instance *>(dev_out);
            m.capacity = dev_cap;
            m.status = d.st();
            TIMED(d, s, VX_K_EMIT, vx_launch_mesh_emit(m, nn, s, &d.launches));
        }
    }
    unsigned char *h_down = d.h_stage + 2048;
    CK(cudaMemcpyAsync(h_down, d_off, down, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h_down + down, d.st(), 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));                     // the one wait
    uint32_t st = 0;
    std::memcpy(&st, h_down + down, 4);
    std::memcpy(offsets, h_down, off_bytes);
    std::memcpy(counts, h_down + off_bytes, 4 * n);
    if (st & VX_ST_BAD_BLOCK)
        return fail(VX_ERR_BAD_BLOCK, "blocks contain an id outside {0 GRASS, 1 DIRT, 2 AIR} (cube.rs:16-21)");
    const uint64_t total = offsets[n];
    if (total_out) *total_out = total;
    if (!instances) return VX_OK;
    if (total > capacity)
        return fail(VX_ERR_CAPACITY, "instance capacity %zu < %llu quads", capacity, (unsigned long long)total);
    if (!dev_out || (st & VX_ST_OVERFLOW)) {
        // does not fit the staging buffer the speculative emit was given: sliced path through the arena
        vx_mesh_args m = emit_args(d, d_coords, d.typ.as<uint32_t>(), d_mm, d_cnt, d_off, nn);
        RET(dev_stream_instances(d, m, offsets, nn, instances));
        return sync_dev(d);
    }
    if (staged) std::memcpy(instances, d.h_inst, total * rec);
    return VX_OK;
}

int vx_mesh(vx_ctx *ctx, const int32_t *coords, size_t n, const uint8_t *blocks, const int32_t *nbr_idx,
            const uint8_t *mesh_mask, uint32_t *counts, uint64_t *offsets, vx_instance *instances, size_t capacity,
            uint64_t *total_out)
{
    if (!ctx) return fail(VX_ERR_BAD_ARG, "ctx is NULL");
    if (total_out) *total_out = 0;
    if (!offsets || (n && (!coords || !blocks || !counts))) return fail(VX_ERR_BAD_ARG, "NULL argument");
    if (n > (size_t)0x3ffffff) return fail(VX_ERR_BAD_ARG, "n too large for one call (%zu)", n);
    offsets[0] = 0;
    if (n == 0) return VX_OK;
    if (nbr_idx)        // validated before anything is enqueued
        for (size_t i = 0; i < n * 4; ++i)
            if (nbr_idx[i] >= (int64_t)n) return fail(VX_ERR_BAD_ARG, "nbr_idx[%zu] = %d out of range", i, nbr_idx[i]);
    if (ctx->devs.size() > 1 && n > kShortList)
        return mesh_multi(ctx, coords, n, blocks, nbr_idx, mesh_mask, counts, offsets, instances, capacity, total_out);
    DevGuard guard(ctx, n);
    if (!guard.d) return cuda_fail(guard.err, "lane setup", __LINE__);
    Dev &d = *guard.d;
    if (n <= kShortList)
        return mesh_short(d, coords, n, blocks, nbr_idx, mesh_mask, counts, offsets,
                          reinterpret_cast<unsigned char *>(instances), capacity, total_out);
    CK(cudaSetDevice(d.id));
    cudaStream_t s = d.stream;
    const uint32_t nn = (uint32_t)n;
    CK(d.coords.ensure(n * 12));
    CK(d.blocks.ensure(n * VX_VOXELS));
    CK(d.counts.ensure(n * 4));
    CK(d.offsets.ensure((n + 1) * 8));
    CK(cudaMemcpyAsync(d.coords.p, coords, n * 12, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d.blocks.p, blocks, n * VX_VOXELS, cudaMemcpyHostToDevice, s));
    const int32_t *d_nbr_idx = nullptr;
    if (nbr_idx) {
        CK(d.nbr_in.ensure(n * 16));
        CK(cudaMemcpyAsync(d.nbr_in.p, nbr_idx, n * 16, cudaMemcpyHostToDevice, s));
        d_nbr_idx = d.nbr_in.as<int32_t>();
    }
    const uint8_t *d_mm = nullptr;
    if (mesh_mask) {
        CK(d.mesh_mask.ensure(n));
        CK(cudaMemcpyAsync(d.mesh_mask.p, mesh_mask, n, cudaMemcpyHostToDevice, s));
        d_mm = d.mesh_mask.as<uint8_t>();
    }
    CK(cudaMemsetAsync(d.st(), 0, 4, s));
    RET(mesh_prepare(d, s, d.coords.as<int32_t>(), nn, d.blocks.as<uint8_t>(), d_nbr_idx, d.counts.as<uint32_t>(),
                     d.offsets.as<uint64_t>(), d_mm, d.st()));
    CK(cudaMemcpyAsync(counts, d.counts.p, n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(offsets, d.offsets.p, (n + 1) * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(d.h_small, d.st(), 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (d.h_small[0] & VX_ST_BAD_BLOCK)
        return fail(VX_ERR_BAD_BLOCK, "blocks contain an id outside {0 GRASS, 1 DIRT, 2 AIR} (cube.rs:16-21)");
    const uint64_t total = offsets[n];
    if (total_out) *total_out = total;
    if (!instances) return VX_OK;
    if (total > capacity)
        return fail(VX_ERR_CAPACITY, "instance capacity %zu < %llu quads", capacity, (unsigned long long)total);
    vx_mesh_args m = emit_args(d, d.coords.as<int32_t>(), d.typ.as<uint32_t>(), d_mm, d.counts.as<uint32_t>(),
                               d.offsets.as<uint64_t>(), nn);
    RET(dev_stream_instances(d, m, offsets, nn, instances));
    RET(sync_dev(d));
    return VX_OK;
}

// ---------------------------------------------------------------------------
// fused generate + mesh
// ---------------------------------------------------------------------------
static int gen_mesh_device_impl(vx_ctx *ctx, const int32_t *coords, size_t n, const vx_noise_params *params,
                                const uint8_t *nbr_mask, uint32_t *counts, uint64_t *offsets, void *instances,
                                size_t capacity, uint32_t *status, uint8_t *blocks_out, void *stream, bool packed,
                                const vx_terrain_opts *opts = nullptr)
{
    if (!ctx) return fail(VX_ERR_BAD_ARG, "ctx is NULL");
    if (!offsets || (n && (!coords || !counts))) return fail(VX_ERR_BAD_ARG, "NULL argument");
    if (n > (size_t)0x3ffffff) return fail(VX_ERR_BAD_ARG, "n too large for one call (%zu)", n);
    vx_noise_params p;
    RET(check_params(params, &p));
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dev &d = ctx->devs[0];
    CK(cudaSetDevice(d.id));
    cudaStream_t s = stream ? (cudaStream_t)stream : d.stream;
    if (status) CK(cudaMemsetAsync(status, 0, 4, s));
    if (n == 0) {
        CK(cudaMemsetAsync(offsets, 0, 8, s));
        return VX_OK;
    }
    const uint32_t nn = (uint32_t)n;
    d.last_n = 0;
    RET(dev_gen_count(d, s, coords, nn, 0, nn, p, nbr_mask, counts, offsets, blocks_out, opts));
    d.last_n = n;
    d.last_typ = false;
    if (instances) {
        vx_mesh_args m = emit_args(d, coords, nullptr, nullptr, counts, offsets, nn);
        m.out = static_cast<vx_instance *>(instances);
        m.packed = packed ? 1 : 0;
        m.capacity = capacity;
        m.status = status;
        TIMED(d, s, VX_K_EMIT, vx_launch_mesh_emit(m, nn, s, &d.launches));
    }
    return VX_OK;
}

int vx_gen_mesh_device(vx_ctx *ctx, const int32_t *coords, size_t n, const vx_noise_params *params,
                       const uint8_t *nbr_mask, uint32_t *counts, uint64_t *offsets, vx_instance *instances,
                       size_t capacity, uint32_t *status, uint8_t *blocks_out, void *stream)
{
    return gen_mesh_device_impl(ctx, coords, n, params, nbr_mask, counts, offsets, instances, capacity, status,
                                blocks_out, stream, false);
}

int vx_emit_again_device(vx_ctx *ctx, const int32_t *coords, size_t n, uint32_t *counts, uint64_t *offsets,
                         vx_instance *instances, size_t capacity, uint32_t *status, void *stream)
{
    if (!ctx) return fail(VX_ERR_BAD_ARG, "ctx is NULL");
    if (n == 0) return VX_OK;
    if (!coords || !counts || !offsets || !instances) return fail(VX_ERR_BAD_ARG, "NULL argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dev &d = ctx->devs[0];
    if (d.last_n != n) return fail(VX_ERR_BAD_ARG, "the context's last device batch had %zu chunks, not %zu", d.last_n, n);
    CK(cudaSetDevice(d.id));
    cudaStream_t s = stream ? (cudaStream_t)stream : d.stream;
    if (status) CK(cudaMemsetAsync(status, 0, 4, s));
    vx_mesh_args m = emit_args(d, coords, d.last_typ ? d.typ.as<uint32_t>() : nullptr, nullptr, counts, offsets, (uint32_t)n);
    m.out = instances;
    m.capacity = capacity;
    m.status = status;
    TIMED(d, s, VX_K_EMIT, vx_launch_mesh_emit(m, (uint32_t)n, s, &d.launches));
    return VX_OK;
}

int vx_gen_mesh_ex_device(vx_ctx *ctx, const int32_t *coords, size_t n, const vx_noise_params *params,
                          const vx_terrain_opts *opts, const uint8_t *nbr_mask, uint32_t *counts, uint64_t *offsets,
                          vx_instance *instances, size_t capacity, uint32_t *status, uint8_t *blocks_out, void *stream)
{
    return gen_mesh_device_impl(ctx, coords, n, params, nbr_mask, counts, offsets, instances, capacity, status,
                                blocks_out, stream, false, opts);
}

int vx_gen_mesh_packed_device(vx_ctx *ctx, const int32_t *coords, size_t n, const vx_noise_params *params,
                              const uint8_t *nbr_mask, uint32_t *counts, uint64_t *offsets, uint32_t *quads,
                              size_t capacity, uint32_t *status, void *stream)
{
    return gen_mesh_device_impl(ctx, coords, n, params, nbr_mask, counts, offsets, quads, capacity, status, nullptr,
                                stream, true);
}

// ---------------------------------------------------------------------------
// Short lists (the reference's one-chunk-per-call site, chunk_builder.rs:80-84): the call is latency-
// bound, so everything is issued before the host waits ONCE.  Coordinates and masks go up in one copy
// from pinned staging; the emit is launched speculatively behind the count (its capacity check runs on
// the device); counts, offsets and the status word come back in one pinned copy.  If the caller's
// instance buffer is pinned host memory (vx_host_alloc) the emit kernel writes straight into it over
// PCIe and no second copy or synchronisation exists at all; otherwise the instances go through the
// arena and one more copy of exactly `total` records.
// ---------------------------------------------------------------------------
static int gen_mesh_short(Dev &d, const int32_t *coords, size_t n, const vx_noise_params &p,
                          const uint8_t *nbr_mask, uint32_t *counts, uint64_t *offsets, unsigned char *instances,
                          size_t capacity, uint64_t *total_out, uint8_t *blocks_out, bool packed,
                          const vx_terrain_opts *opts)
{
    const size_t rec = packed ? sizeof(uint32_t) : sizeof(vx_instance);
    const uint32_t nn = (uint32_t)n;
    CK(cudaSetDevice(d.id));
    cudaStream_t s = d.stream;
    // upload: [coords 12 n][mask n]
    const size_t up = 12 * n + (nbr_mask ? n : 0);
    CK(d.coords.ensure(up));
    std::memcpy(d.h_stage, coords, 12 * n);
    if (nbr_mask) std::memcpy(d.h_stage + 12 * n, nbr_mask, n);
    CK(cudaMemcpyAsync(d.coords.p, d.h_stage, up, cudaMemcpyHostToDevice, s));
    const uint8_t *d_mask = nbr_mask ? d.coords.as<uint8_t>() + 12 * n : nullptr;
    // download block on the device: [offsets 8 (n + 1)][counts 4 n], contiguous so ONE copy brings both back
    const size_t off_bytes = 8 * (n + 1), down = off_bytes + 4 * n;
    CK(d.offsets.ensure(down));
    uint64_t *d_off = d.offsets.as<uint64_t>();
    uint32_t *d_cnt = reinterpret_cast<uint32_t *>(d.offsets.as<unsigned char>() + off_bytes);
    uint8_t *d_blocks = nullptr;
    if (blocks_out) {
        CK(d.blocks.ensure(n * VX_VOXELS));
        d_blocks = d.blocks.as<uint8_t>();
    }
    CK(cudaMemsetAsync(d.st(), 0, 4, s));
    RET(dev_gen_count(d, s, d.coords.as<int32_t>(), nn, 0, nn, p, d_mask, d_cnt, d_off, d_blocks, opts));
    // where the instances go: straight into a pinned caller buffer, else the arena
    void *dev_out = nullptr;
    size_t dev_cap = 0;
    bool direct = false, staged = false;
    if (instances) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, instances) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
            dev_out = at.devicePointer;
            dev_cap = capacity;
            direct = true;
        } else {
            cudaGetLastError();                       // a pageable pointer is not an error here
            // Pageable destination: a cudaMemcpy into it would go through the driver's own staging, one caller at
            // a time.  The emit writes into this lane's pinned buffer over PCIe instead and the CALLING thread copies
            // from there (pool threads copy in parallel); a list that does not fit falls back to the arena below.
            // (terrain has at most ~8 k quads per chunk: longer lists would overflow the buffer and emit twice)
            const bool fits = n * (size_t)8192 * rec <= kHostInstBytes;
            if (fits && !d.h_inst && cudaHostAlloc((void **)&d.h_inst, kHostInstBytes, cudaHostAllocPortable) != cudaSuccess) {
                cudaGetLastError();
                d.h_inst = nullptr;
            }
            void *dp = nullptr;
            if (fits && d.h_inst && cudaHostGetDevicePointer(&dp, d.h_inst, 0) == cudaSuccess) {
                dev_out = dp;
                dev_cap = std::min(capacity, kHostInstBytes / rec);
                staged = true;
            } else {
                cudaGetLastError();
                // a lane's arena only has to hold this list (n <= kShortList chunks)
                const size_t arena = std::min(kArenaBytes, n * (size_t)VX_MAX_QUADS_PER_CHUNK * rec);
                CK(d.inst[0].ensure(arena));
                dev_out = d.inst[0].p;
                dev_cap = std::min(capacity, arena / rec);
            }
        }
        vx_mesh_args m = emit_args(d, d.coords.as<int32_t>(), nullptr, nullptr, d_cnt, d_off, nn);
        m.out = static_cast<vx_instance *>(dev_out);
        m.packed = packed ? 1 : 0;
        m.capacity = dev_cap;
        m.status = d.st();
        TIMED(d, s, VX_K_EMIT, vx_launch_mesh_emit(m, nn, s, &d.launches));
    }
    unsigned char *h_down = d.h_stage + 2048;
    CK(cudaMemcpyAsync(h_down, d_off, down, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h_down + down, d.st(), 4, cudaMemcpyDeviceToHost, s));
    // cube ids: a few chunks go straight to the caller; more of them into pageable memory go through the bounce buffers
    const bool blocks_bounce = blocks_out && use_bounce(blocks_out, n * VX_VOXELS);
    if (blocks_out && !blocks_bounce) CK(cudaMemcpyAsync(blocks_out, d_blocks, n * VX_VOXELS, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));                     // the one wait
    if (blocks_bounce) RET(d2h_pageable(d, s, blocks_out, d_blocks, n * VX_VOXELS));
    std::memcpy(offsets, h_down, off_bytes);
    std::memcpy(counts, h_down + off_bytes, 4 * n);
    uint32_t st = 0;
    std::memcpy(&st, h_down + down, 4);
    const uint64_t total = offsets[n];
    if (total_out) *total_out = total;
    if (!instances) return VX_OK;
    if (total > capacity)
        return fail(VX_ERR_CAPACITY, "instance capacity %zu < %llu quads", capacity, (unsigned long long)total);
    if (st & VX_ST_OVERFLOW) {
        // fits the caller's buffer but not the arena slice the speculative emit was given: sliced path
        vx_mesh_args m = emit_args(d, d.coords.as<int32_t>(), nullptr, nullptr, d_cnt, d_off, nn);
        RET(dev_stream_instances(d, m, offsets, nn, instances, packed));
        return sync_dev(d);
    }
    if (staged) {
        std::memcpy(instances, d.h_inst, total * rec);
    } else if (!direct && total) {
        CK(cudaMemcpyAsync(instances, dev_out, total * rec, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    return VX_OK;
}

static int gen_mesh_host_impl(vx_ctx *ctx, const int32_t *coords, size_t n, const vx_noise_params *params,
                              const uint8_t *nbr_mask, uint32_t *counts, uint64_t *offsets, void *instances_v,
                              size_t capacity, uint64_t *total_out, uint8_t *blocks_out, bool packed,
                              const vx_terrain_opts *opts = nullptr)
{
    unsigned char *instances = static_cast<unsigned char *>(instances_v);
    const size_t rec = packed ? sizeof(uint32_t) : sizeof(vx_instance);
    if (!ctx) return fail(VX_ERR_BAD_ARG, "ctx is NULL");
    if (total_out) *total_out = 0;
    if (!offsets || (n && (!coords || !counts))) return fail(VX_ERR_BAD_ARG, "NULL argument");
    if (n > (size_t)0x3ffffff) return fail(VX_ERR_BAD_ARG, "n too large for one call (%zu)", n);
    offsets[0] = 0;
    if (n == 0) return VX_OK;
    vx_noise_params p;
    RET(check_params(params, &p));
    if (n <= kShortList && !ctx->devs[0].timing) {
        // concurrent callers each lease a lane (own scratch + streams); `mu` is not taken
        LaneLease lease(ctx);
        if (!lease.d) return cuda_fail(lease.err, "lane setup", __LINE__);
        return gen_mesh_short(*lease.d, coords, n, p, nbr_mask, counts, offsets, instances, capacity, total_out,
                              blocks_out, packed, opts);
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t nd = ctx->devs.size();
    if (n <= kShortList)
        return gen_mesh_short(ctx->devs[0], coords, n, p, nbr_mask, counts, offsets, instances, capacity, total_out,
                              blocks_out, packed, opts);
    const uint32_t nn = (uint32_t)n;
    std::vector<uint32_t> lo(nd), hi(nd);
    std::vector<std::vector<uint64_t>> loc(nd);     // device-local offsets

    RET(for_each_device(nd, [&](size_t di) -> int {
        Dev &d = ctx->devs[di];
        split_range(n, nd, di, 1, &lo[di], &hi[di]);
        const uint32_t c = hi[di] - lo[di];
        if (c == 0) return VX_OK;
        CK(cudaSetDevice(d.id));
        CK(d.coords.ensure(n * 12));
        CK(cudaMemcpyAsync(d.coords.p, coords, n * 12, cudaMemcpyHostToDevice, d.stream));
        const uint8_t *d_mask = nullptr;
        if (nbr_mask) {
            CK(d.mask8.ensure(n));
            CK(cudaMemcpyAsync(d.mask8.p, nbr_mask, n, cudaMemcpyHostToDevice, d.stream));
            d_mask = d.mask8.as<uint8_t>();
        }
        CK(d.counts.ensure((size_t)c * 4));
        CK(d.offsets.ensure(((size_t)c + 1) * 8));
        uint8_t *d_blocks = nullptr;
        if (blocks_out) {
            CK(d.blocks.ensure((size_t)c * VX_VOXELS));
            d_blocks = d.blocks.as<uint8_t>();
        }
        CK(cudaMemsetAsync(d.st(), 0, 4, d.stream));
        RET(dev_gen_count(d, d.stream, d.coords.as<int32_t>(), nn, lo[di], hi[di], p, d_mask,
                          d.counts.as<uint32_t>(), d.offsets.as<uint64_t>(), d_blocks, opts));
        loc[di].resize((size_t)c + 1);
        CK(cudaMemcpyAsync(counts + lo[di], d.counts.p, (size_t)c * 4, cudaMemcpyDeviceToHost, d.stream));
        CK(cudaMemcpyAsync(loc[di].data(), d.offsets.p, ((size_t)c + 1) * 8, cudaMemcpyDeviceToHost, d.stream));
        const bool blocks_bounce = blocks_out && use_bounce(blocks_out + (size_t)lo[di] * VX_VOXELS, (size_t)c * VX_VOXELS);
        if (blocks_out && !blocks_bounce)
            CK(cudaMemcpyAsync(blocks_out + (size_t)lo[di] * VX_VOXELS, d.blocks.p, (size_t)c * VX_VOXELS,
                               cudaMemcpyDeviceToHost, d.stream));
        CK(cudaStreamSynchronize(d.stream));
        if (blocks_bounce)      // pageable: through the pinned bounce buffers
            RET(d2h_pageable(d, d.stream, blocks_out + (size_t)lo[di] * VX_VOXELS, d.blocks.p, (size_t)c * VX_VOXELS));
        return VX_OK;
    }));
    uint64_t total = 0;
    std::vector<uint64_t> base(nd, 0);
    for (size_t di = 0; di < nd; ++di) {
        if (hi[di] == lo[di]) continue;
        base[di] = total;
        const uint32_t c = hi[di] - lo[di];
        for (uint32_t i = 0; i < c; ++i) offsets[lo[di] + i] = total + loc[di][i];
        total += loc[di][c];
    }
    offsets[n] = total;
    if (total_out) *total_out = total;
    if (!instances) return VX_OK;
    if (total > capacity)
        return fail(VX_ERR_CAPACITY, "instance capacity %zu < %llu quads", capacity, (unsigned long long)total);
    return for_each_device(nd, [&](size_t di) -> int {
        const uint32_t c = hi[di] - lo[di];
        if (c == 0) return VX_OK;
        Dev &d = ctx->devs[di];
        CK(cudaSetDevice(d.id));
        vx_mesh_args m = emit_args(d, d.coords.as<int32_t>() + 3 * (size_t)lo[di], nullptr, nullptr,
                                   d.counts.as<uint32_t>(), d.offsets.as<uint64_t>(), c);
        RET(dev_stream_instances(d, m, loc[di].data(), c, instances + base[di] * rec, packed));
        return sync_dev(d);
    });
}

int vx_gen_mesh(vx_ctx *ctx, const int32_t *coords, size_t n, const vx_noise_params *params,
                const uint8_t *nbr_mask, uint32_t *counts, uint64_t *offsets, vx_instance *instances,
                size_t capacity, uint64_t *total_out, uint8_t *blocks_out)
{
    return gen_mesh_host_impl(ctx, coords, n, params, nbr_mask, counts, offsets, instances, capacity, total_out,
                              blocks_out, false);
}

int vx_gen_mesh_ex(vx_ctx *ctx, const int32_t *coords, size_t n, const vx_noise_params *params,
                   const vx_terrain_opts *opts, const uint8_t *nbr_mask, uint32_t *counts, uint64_t *offsets,
                   vx_instance *instances, size_t capacity, uint64_t *total_out, uint8_t *blocks_out)
{
    return gen_mesh_host_impl(ctx, coords, n, params, nbr_mask, counts, offsets, instances, capacity, total_out,
                              blocks_out, false, opts);
}

int vx_gen_mesh_packed(vx_ctx *ctx, const int32_t *coords, size_t n, const vx_noise_params *params,
                       const uint8_t *nbr_mask, uint32_t *counts, uint64_t *offsets, uint32_t *quads,
                       size_t capacity, uint64_t *total_out)
{
    return gen_mesh_host_impl(ctx, coords, n, params, nbr_mask, counts, offsets, quads, capacity, total_out, nullptr,
                              true);
}

// ---------------------------------------------------------------------------
// frustum
// ---------------------------------------------------------------------------
int vx_planes_from_matrix(const float m[16], float planes[24])
{
    if (!m || !planes) return fail(VX_ERR_BAD_ARG, "NULL argument");
    // frustum_culling.rs:47-85: data[c][r] = m[c*4 + r]; planes = row3 +- row{0,1,2};
    // stored order near, far, left, right, top, bottom (:84); not normalised.
    // (24 scalar adds, once per frame -- host-side bookkeeping, not the data path.)
    for (int c = 0; c < 4; ++c) {
        const float r0 = m[c * 4 + 0], r1 = m[c * 4 + 1], r2 = m[c * 4 + 2], r3 = m[c * 4 + 3];
        planes[0 * 4 + c] = r3 + r2;   // near   :75-78
        planes[1 * 4 + c] = r3 - r2;   // far    :80-83
        planes[2 * 4 + c] = r3 + r0;   // left   :55-58
        planes[3 * 4 + c] = r3 - r0;   // right  :60-63
        planes[4 * 4 + c] = r3 - r1;   // top    :65-68
        planes[5 * 4 + c] = r3 + r1;   // bottom :70-73
    }
    return VX_OK;
}

int vx_frustum_cull_device(vx_ctx *ctx, const float *planes, size_t n_frames, const int32_t *coords, size_t n,
                           uint8_t *visible, uint32_t *bits, void *stream)
{
    if (!ctx) return fail(VX_ERR_BAD_ARG, "ctx is NULL");
    if (n == 0 || n_frames == 0) return VX_OK;
    if (!planes || !coords || (!visible && !bits)) return fail(VX_ERR_BAD_ARG, "NULL argument");
    if (n > 0x7fffffffu || n_frames > 65535u * 64u) return fail(VX_ERR_BAD_ARG, "n or n_frames too large");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dev &d = ctx->devs[0];
    CK(cudaSetDevice(d.id));
    cudaStream_t s = stream ? (cudaStream_t)stream : d.stream;
    TIMED(d, s, VX_K_FRUSTUM, vx_launch_frustum(planes, (uint32_t)n_frames, coords, (uint32_t)n, visible, bits, s, &d.launches));
    return VX_OK;
}

static int frustum_host(vx_ctx *ctx, const float *planes, size_t n_frames, const int32_t *coords, size_t n,
                        uint8_t *visible, uint32_t *bits)
{
    if (!ctx) return fail(VX_ERR_BAD_ARG, "ctx is NULL");
    if (n == 0 || n_frames == 0) return VX_OK;
    if (!planes || !coords || (!visible && !bits)) return fail(VX_ERR_BAD_ARG, "NULL argument");
    if (n > 0x7fffffffu) return fail(VX_ERR_BAD_ARG, "n too large");
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t nd = ctx->devs.size();
    const size_t n_words = (n + 31) / 32;
    for (size_t di = 0; di < nd; ++di) {
        Dev &d = ctx->devs[di];
        uint32_t lo, hi;
        split_range(n, nd, di, 32, &lo, &hi);       // shards start on a bit-word boundary
        if (hi <= lo) continue;
        const size_t c = hi - lo, cw = (c + 31) / 32;
        CK(cudaSetDevice(d.id));
        CK(d.coords.ensure(c * 12));
        CK(cudaMemcpyAsync(d.coords.p, coords + 3 * (size_t)lo, c * 12, cudaMemcpyHostToDevice, d.stream));
        const size_t per_frame = visible ? c : cw * 4;
        // frames per launch: the staging slice, and gridDim.y = ceil(frames / 64) <= 65535
        const size_t fs = std::min({n_frames, std::max<size_t>(1, kSliceBytes / per_frame), (size_t)65535 * 64});
        CK(d.planes.ensure(n_frames * 96));
        CK(cudaMemcpyAsync(d.planes.p, planes, n_frames * 96, cudaMemcpyHostToDevice, d.stream));
        if (visible) CK(d.vis.ensure(fs * c));
        else CK(d.bits.ensure(fs * cw * 4));
        for (size_t f = 0; f < n_frames; f += fs) {
            const size_t nf = std::min(fs, n_frames - f);
            TIMED(d, d.stream, VX_K_FRUSTUM,
                  vx_launch_frustum(d.planes.as<float>() + f * 24, (uint32_t)nf, d.coords.as<int32_t>(), (uint32_t)c,
                                    visible ? d.vis.as<uint8_t>() : nullptr,
                                    visible ? nullptr : d.bits.as<uint32_t>(), d.stream, &d.launches));
            if (visible)
                CK(cudaMemcpy2DAsync(visible + f * n + lo, n, d.vis.p, c, c, nf, cudaMemcpyDeviceToHost, d.stream));
            else
                CK(cudaMemcpy2DAsync(bits + f * n_words + lo / 32, n_words * 4, d.bits.p, cw * 4, cw * 4, nf,
                                     cudaMemcpyDeviceToHost, d.stream));
        }
    }
    for (Dev &d : ctx->devs) RET(sync_dev(d));
    return VX_OK;
}

int vx_frustum_cull(vx_ctx *ctx, const float *planes, size_t n_frames, const int32_t *coords, size_t n,
                    uint8_t *visible)
{
    if (!visible && n && n_frames) return fail(VX_ERR_BAD_ARG, "visible is NULL");
    return frustum_host(ctx, planes, n_frames, coords, n, visible, nullptr);
}

int vx_frustum_cull_bits(vx_ctx *ctx, const float *planes, size_t n_frames, const int32_t *coords, size_t n,
                         uint32_t *bits)
{
    if (!bits && n && n_frames) return fail(VX_ERR_BAD_ARG, "bits is NULL");
    return frustum_host(ctx, planes, n_frames, coords, n, nullptr, bits);
}

// ---------------------------------------------------------------------------
int vx_set_timing(vx_ctx *ctx, int on)
{
    if (!ctx) return fail(VX_ERR_BAD_ARG, "ctx is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (Dev &d : ctx->devs) d.timing = on != 0;
    return VX_OK;
}

int vx_get_timing(vx_ctx *ctx, double ms[VX_K_COUNT_KINDS], uint64_t launches[VX_K_COUNT_KINDS])
{
    if (!ctx || !ms || !launches) return fail(VX_ERR_BAD_ARG, "NULL argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (int k = 0; k < VX_K_COUNT_KINDS; ++k) {
        ms[k] = 0.0;
        launches[k] = 0;
    }
    for (Dev &d : ctx->devs) {
        CK(cudaSetDevice(d.id));
        for (const Dev::Stamp &st : d.stamps) {
            CK(cudaEventSynchronize(st.b));
            float t = 0.f;
            CK(cudaEventElapsedTime(&t, st.a, st.b));
            ms[st.kind] += t;
            launches[st.kind] += 1;
            d.free_events.emplace_back(st.a, st.b);
        }
        d.stamps.clear();
    }
    return VX_OK;
}

int vx_measure_peak(vx_ctx *ctx, int kind, double *tera_ops)
{
    if (!ctx || !tera_ops) return fail(VX_ERR_BAD_ARG, "NULL argument");
    if (kind < 0 || kind > 2) return fail(VX_ERR_BAD_ARG, "kind must be 0, 1 or 2");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Dev &d = ctx->devs[0];
    CK(cudaSetDevice(d.id));
    CK(vx_run_peak(kind, d.n_sm, tera_ops, d.stream, &d.launches));
    return VX_OK;
}

}  // extern "C"
