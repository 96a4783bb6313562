// Engine + planner: lowers a batch of graph descriptions (wae_graph.h) into GPU stages (wae_device.h) and runs
// them chunk by chunk on one CUDA stream.  C ABI: wae_engine_*, wae_batch_*, wae_render_batch (include/wae.h).
//
// Reference functions replaced by this file:
//   RenderThread::render_audiobuffer_sync / render_offline_quantum   src/render/thread.rs:260-302,355-396
//   Graph::order_nodes / visit / render                              src/render/graph.rs:331-591
//   (the per-node arithmetic is in wae_kernels.cu)
#include "wae_graph.h"
#include "wae_hostmath.h"
#include "wae_hrtf_host.h"
#include "wae_resample_host.h"
#include "wae_kernels.h"
#include "wae_param_core.h"
#include "wae_param_host.h"

#include <cuda_runtime.h>
#include <emmintrin.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <fstream>
#include <mutex>
#include <sched.h>
#include <sstream>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <tuple>
#include <unordered_map>
#include <unordered_set>

using namespace wae;
namespace hm = wae::hostmath;

#define CUDA_TRY(expr)                                                                                       \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess) return fail(WAE_CUDA_ERROR, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// host worker threads of an engine: planning of graph groups and the copy-out of rendered PCM to pageable caller memory
struct WorkerPool {
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::function<void()>> q;
    bool stop = false;
    WorkerPool(int n, int device) {
        for (int i = 0; i < n; i++)
            threads.emplace_back([this, device] {
                cudaSetDevice(device);
                for (;;) {
                    std::function<void()> f;
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [this] { return stop || !q.empty(); });
                        if (q.empty()) return;
                        f = std::move(q.front());
                        q.pop_front();
                    }
                    f();
                }
            });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        for (auto& t : threads) t.join();
    }
    void submit(std::function<void()> f) {
        {
            std::lock_guard<std::mutex> lk(mu);
            q.push_back(std::move(f));
        }
        cv.notify_one();
    }
    int size() const { return (int)threads.size(); }
    // fn(i) for i in [0, n), on the workers; returns when all are done
    void parallel_for(int n, const std::function<void(int)>& fn) {
        if (n <= 0) return;
        std::mutex dm;
        std::condition_variable dcv;
        int left = n;
        for (int i = 0; i < n; i++)
            submit([&, i] {
                fn(i);
                std::lock_guard<std::mutex> lk(dm);
                if (--left == 0) dcv.notify_all();
            });
        std::unique_lock<std::mutex> lk(dm);
        dcv.wait(lk, [&] { return left == 0; });
    }
};

struct wae_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;  // copy streams of the pipelined / one-shot paths
    int64_t chunk_frames = 0;  // 0 = auto
    bool fuse = true;
    int voice_sum = -1;  // WAE_OPT_VOICE_SUM: fused oscillator voices + ordered sum (k_voice_sum); -1: WAE_VOICE_SUM from the environment, default on
    bool serial_filters = false;
    int pipeline_groups = 0;  // 0 = auto
    int param_parallel = 2;  // WAE_OPT_PARAM_PARALLEL: 2 k_param_spec (CTA per param, speculative walks of 32 quanta), 1 k_param_parallel (warp per param), 0 k_param (lane 0 evaluates every frame)
    float* d_sine = nullptr;
    wae::HrirSphere* sphere = nullptr;  // wae_engine_set_hrir_sphere
    float* d_sphere_ir = nullptr;
    float* d_sphere_pos = nullptr;
    uint32_t* d_sphere_tri = nullptr;
    struct RateSphere {  // the sphere's responses resampled to a context rate (HrirSphere::new of the crate), built on first use
        float* d_ir = nullptr;
        uint32_t taps = 0;
        std::vector<float> ir_host;  // [vertex][2][taps] (static panners: blended on the host into the response of a convolution)
    };
    std::map<uint32_t, RateSphere> sphere_rates;
    std::mutex sphere_mu;
    void drop_rate_spheres() {
        for (auto& kv : sphere_rates)
            if (kv.second.d_ir) cudaFree(kv.second.d_ir);
        sphere_rates.clear();
    }
    // ---- device memory of finished batches is kept and handed to the next batch (a render call that prepares, renders and drops
    // its batch would otherwise pay cudaMalloc / cudaFree — both synchronising — for gigabytes of PCM every time)
    std::mutex mem_mu;
    std::multimap<size_t, void*> dev_free;          // cached blocks by size
    std::unordered_map<void*, size_t> dev_size;     // every live block (handed out or cached)
    size_t dev_cached_bytes = 0;
    static size_t round_block(size_t b) {
        if (b < 512) return 512;
        if (b < ((size_t)1 << 16)) return (b + 511) / 512 * 512;
        if (b < ((size_t)2 << 20)) return (b + 65535) / 65536 * 65536;
        return (b + (((size_t)2 << 20) - 1)) / ((size_t)2 << 20) * ((size_t)2 << 20);
    }
    void* dev_alloc(size_t bytes, bool* fresh = nullptr) {
        const size_t r = round_block(bytes);
        {
            std::lock_guard<std::mutex> lk(mem_mu);
            auto it = dev_free.lower_bound(r);
            if (it != dev_free.end() && it->first <= r + std::max<size_t>(r / 8, 4096)) {
                void* p = it->second;
                dev_cached_bytes -= it->first;
                dev_free.erase(it);
                if (fresh) *fresh = false;
                return p;
            }
        }
        void* p = nullptr;
        if (cudaMalloc(&p, r) != cudaSuccess) {
            cudaGetLastError();
            dev_trim();  // give the cache back and try once more
            if (cudaMalloc(&p, r) != cudaSuccess) {
                cudaGetLastError();
                return nullptr;
            }
        }
        std::lock_guard<std::mutex> lk(mem_mu);
        dev_size[p] = r;
        if (fresh) *fresh = true;
        return p;
    }
    void dev_release(void* p) {
        std::lock_guard<std::mutex> lk(mem_mu);
        auto it = dev_size.find(p);
        if (it == dev_size.end()) return;
        dev_free.emplace(it->second, p);
        dev_cached_bytes += it->second;
    }
    void dev_trim() {
        std::lock_guard<std::mutex> lk(mem_mu);
        for (auto& kv : dev_free) {
            cudaFree(kv.second);
            dev_size.erase(kv.second);
        }
        dev_free.clear();
        dev_cached_bytes = 0;
    }
    // ---- host side of the one-shot render: worker threads, page-locked staging slots for pageable output buffers
    WorkerPool* pool = nullptr;
    int n_workers = 0;  // 0 = auto
    WorkerPool* workers() {
        if (!pool) {
            int n = n_workers;
            if (n <= 0) {
                const unsigned hw = std::thread::hardware_concurrency();
                n = (int)std::min<unsigned>(16u, std::max<unsigned>(2u, hw / 8u));
            }
            pool = new WorkerPool(n, device);
        }
        return pool;
    }
    static constexpr int kStageSlots = 4;
    float* h_stage[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
    size_t h_stage_bytes = 0;
    bool ensure_stage(size_t bytes) {
        if (h_stage_bytes >= bytes) return true;
        for (int i = 0; i < kStageSlots; i++) {
            if (h_stage[i]) cudaFreeHost(h_stage[i]);
            h_stage[i] = nullptr;
        }
        h_stage_bytes = 0;
        for (int i = 0; i < kStageSlots; i++)
            if (cudaHostAlloc((void**)&h_stage[i], bytes, cudaHostAllocDefault) != cudaSuccess) {
                cudaGetLastError();
                return false;
            }
        h_stage_bytes = bytes;
        return true;
    }
    // small ring of page-locked pieces for the copy-out to pageable caller memory (render_oneshot_host)
    char* h_ring = nullptr;
    size_t h_ring_bytes = 0;
    bool ensure_ring(size_t bytes) {
        if (h_ring_bytes >= bytes) return true;
        if (h_ring) cudaFreeHost(h_ring);
        h_ring = nullptr;
        h_ring_bytes = 0;
        if (cudaHostAlloc((void**)&h_ring, bytes, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        h_ring_bytes = bytes;
        return true;
    }
    std::string numa_cpus;  // CPUs this engine's host threads were bound to (WAE_OPT_BIND_NUMA), for the record
};

namespace wae {
int engine_device(const wae_engine* eng) { return eng ? eng->device : 0; }
}  // namespace wae

namespace {

// (the order of the kinds is the launch order inside one level: mixes first; k_delay_mono before the delay reader that needs it)
enum StageKind : int {
    S_MIX = 0, S_MIX_DYN, S_OSC, S_CONST, S_ABSN, S_BIQUAD, S_IIR, S_GAIN, S_SHAPER, S_SPAN, S_PAN, S_ROUTE, S_DELAY_MONO, S_DELAY, S_DELAY_WRITE, S_COMP, S_ANALYSER,
    S_CONV_FFT, S_CONV_MAC, S_CONV_MAC_ACC, S_CHAIN, S_PARAM, S_OSC_AR, S_BIQUAD_AR, S_ABSN_SLOW, S_HRTF, S_PAN_DYN, S_ABSN_SERIAL, S_SHAPER_OS, S_META, S_VSUM, S_KINDS
};
const char* kStageNames[S_KINDS] = {"k_mix", "k_mix_dyn", "k_oscillator", "k_constant", "k_buffer_source", "k_biquad_serial", "k_iir_serial", "k_gain",
                                    "k_shaper", "k_stereo_panner", "k_panner_eq", "k_route", "k_delay_mono", "k_delay_read", "k_ring_write", "k_compressor",
                                    "k_analyser", "k_conv_fft_in", "k_conv_mac_ifft", "k_conv_mac_ifft(acc)", "k_chain", "k_param", "k_osc_arate", "k_biquad_arate", "k_buffer_source_slow", "k_hrtf_fir", "k_panner_dyn", "k_buffer_source_serial", "k_shaper_os", "k_meta", "k_voice_sum"};

// host-side accumulation of instances for one (level, kind) stage
struct StageBuild {
    int cls = 0;
    int level = 0;
    int kind = 0;
    int variant = 0;
    std::vector<OscInst> osc;
    std::vector<ConstInst> cst;
    std::vector<AbsnInst> absn;
    std::vector<BiquadInst> biquad;
    std::vector<ChainInst> chain;
    std::vector<ParamInst> param;
    std::vector<OscArInst> osc_ar;
    std::vector<BiquadArInst> biquad_ar;
    std::vector<AbsnSlowInst> absn_slow;
    std::vector<ScanCoef> scan_coef;
    size_t n_scan_coef = 0;  // sets appended (the sizing pass counts them without building them)
    // one set of scan constants per biquad of a chain instance, in instance order; returns its index
    template <typename MakeFn>
    int32_t add_scan_coef(bool build, MakeFn&& make) {
        if (build) scan_coef.push_back(make());
        return (int32_t)n_scan_coef++;
    }
    std::vector<IirInst> iir;
    std::vector<GainInst> gain;
    std::vector<ShaperInst> shaper;
    std::vector<SPanInst> span;
    std::vector<float2> span_gains;
    std::vector<PanInst> pan;
    std::vector<HrtfInst> hrtf;
    std::vector<HrtfSelInst> hrtf_sel;
    std::vector<PanDynInst> pan_dyn;
    std::vector<AbsnSerialInst> absn_serial;
    std::vector<ShaperOsInst> shaper_os;
    std::vector<RouteInst> route;
    std::vector<DelayInst> delay;
    std::vector<CompInst> comp;
    std::vector<AnalyserInst> analyser;
    std::vector<MixInst> mix;
    std::vector<MixEdge> mix_edges;
    std::vector<MixDynInst> mix_dyn;
    std::vector<MetaInst> meta;
    std::vector<ConvInput> conv_in;
    std::vector<ConvPath> conv_path;
    std::vector<VoiceGroup> vgroups;  // S_VSUM: groups of consecutive `chain` records
    int max_ch = 1;
};

struct Stage {
    int cls = 0;  // see Planner::stage()
    int seg = 0;  // render segment (between two suspend points) this stage belongs to
    int kind = 0;
    int variant = 0;
    int group = 0;
    int n = 0;
    int n_b = 0;
    int max_ch = 1;
    void* d_a = nullptr;  // instances
    void* d_b = nullptr;  // auxiliary table (mix edges, scan coefficients, conv inputs, panner gains)
    void* d_c = nullptr;  // S_VSUM: voice groups
    float ms = 0.f;       // accumulated device time of the last run (when timing is enabled)
    ChainAux chain;       // S_CHAIN with biquads: ticket counter + slab hand-off slots (k_chain)
};

struct AnalyserRec {
    uint32_t graph_index;
    uint32_t node;
    float* d_ring;
    uint32_t fft_size;
    double smoothing;
    float* d_last_fft;  // last_fft_output (analysis.rs:168), zeroed per run
    float* d_db;        // read-out scratch
    bool computed;      // frequency data already computed for the end-of-render time (analysis.rs:353-361)
    double min_db, max_db;
};

}  // namespace

struct wae_batch {
    wae_engine* engine = nullptr;
    uint32_t n_graphs = 0, channels = 0;
    uint64_t length = 0;   // frames requested
    int64_t lq = 0;        // frames rendered: whole quanta (src/render/thread.rs:273)
    int64_t chunk = 0;     // frames per chunk
    std::vector<void*> allocs;
    std::vector<Stage> stages;
    float* d_out = nullptr;  // [n_graphs][channels][length]
    // state that must be reset before every run
    std::vector<std::pair<void*, size_t>> zero_on_run;
    std::vector<AnalyserRec> analysers;
    struct CompRec { uint32_t graph; wae_node_id node; const float* d_state; };
    std::vector<CompRec> compressors;
    // source PCM assets: device destination <- host source (re-uploadable: wae_batch_upload)
    // Graph groups: the batch is cut into contiguous groups of graphs; a group's source PCM lives in one device slab
    // mirrored by one pinned host slab, so that H2D(group k+1), render(group k) and D2H(group k-1) overlap on three
    // streams (wae_batch_run_pipelined).  All groups share the arena-sizing chunk.
    struct Group {
        uint32_t g0 = 0, g1 = 0;        // graphs [g0, g1)
        size_t stage0 = 0, stage1 = 0;  // stages [stage0, stage1) of `stages` (all segments)
        std::vector<std::pair<size_t, size_t>> seg_stages;  // per render segment: its stages
        std::vector<int64_t> seg_bounds;                    // 0 = b0 < b1 < ... < lq: the suspend frames of this group's graphs
        float* d_src = nullptr;         // device slab of source PCM
        float* h_src = nullptr;         // pinned host mirror, built on first use (wae_batch_upload / wae_batch_run_pipelined)
        struct SrcCopy {                // the PCM of one AudioBufferSourceNode inside the slab: planar [ch][stride], as PcmBuffer holds it
            std::shared_ptr<PcmBuffer> buf;
            size_t offset, floats;      // floats
        };
        std::vector<SrcCopy> src_copies;
        std::vector<size_t> graph_src_base;  // groups without suspend points: the slab cursor each graph starts at (split planning)
        size_t src_floats = 0;
        cudaEvent_t ev_h2d = nullptr, ev_done = nullptr;
    };
    std::vector<Group> groups;
    // OfflineAudioContext::suspend_sync: a group's render is cut at the suspend frames of its graphs (graphs with different
    // suspend points are put in different groups); every segment has its own plan, node state is shared between the plans
    // through `state_map` (graph, node, allocation sequence, salt)
    struct StateKey {
        uint32_t graph, node, seq;
        uint64_t salt;
        bool operator<(const StateKey& o) const { return std::tie(graph, node, seq, salt) < std::tie(o.graph, o.node, o.seq, o.salt); }
    };
    std::map<StateKey, std::pair<void*, size_t>> state_map;
    std::vector<void*> pinned;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;  // the engine's copy streams
    std::recursive_mutex mu;  // groups are planned on worker threads: allocation, state map, read-out records
    struct Timed {
        size_t stage, e0, e1;
    };
    std::vector<Timed> timed;
    wae_batch_stats stats{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> stage_events;
    bool time_stages = false;
    size_t timed_events_used = 0;
    std::atomic<uint64_t> arena_bytes{0}, asset_bytes{0};
    uint64_t n_cuda_malloc = 0;  // prepare-time diagnostics (WAE_PREPARE_PROFILE=1): blocks that were not served from the engine's cache

    // small per-node state that is zeroed before every run lives in slabs: one memset per slab, not per node
    char* slab = nullptr;
    size_t slab_used = 0, slab_cap = 0;
    void* slab_alloc(size_t bytes) {
        bytes = (bytes + 255) / 256 * 256;
        if (bytes > (1u << 20)) return nullptr;
        if (!slab || slab_used + bytes > slab_cap) {
            slab_cap = 8u << 20;
            bool fresh = false;
            void* p = engine->dev_alloc(slab_cap, &fresh);
            if (!p) return nullptr;
            n_cuda_malloc += fresh ? 1 : 0;
            allocs.push_back(p);
            cudaMemsetAsync(p, 0, slab_cap, engine->stream);
            zero_on_run.push_back({p, slab_cap});
            slab = (char*)p;
            slab_used = 0;
        }
        void* r = slab + slab_used;
        slab_used += bytes;
        return r;
    }
    // small uploads / tables share slabs too (one device block per 4 MiB instead of one per table)
    char* tslab = nullptr;
    size_t tslab_used = 0, tslab_cap = 0;
    void* table_alloc(size_t bytes) {
        bytes = (bytes + 255) / 256 * 256;
        if (bytes > (512u << 10)) return nullptr;
        if (!tslab || tslab_used + bytes > tslab_cap) {
            tslab_cap = 4u << 20;
            bool fresh = false;
            void* p = engine->dev_alloc(tslab_cap, &fresh);
            if (!p) return nullptr;
            n_cuda_malloc += fresh ? 1 : 0;
            allocs.push_back(p);
            tslab = (char*)p;
            tslab_used = 0;
        }
        void* r = tslab + tslab_used;
        tslab_used += bytes;
        return r;
    }
    template <typename T>
    T* dalloc(size_t count, bool zero = false, bool rezero_on_run = false) {
        std::lock_guard<std::recursive_mutex> lk(mu);
        if (rezero_on_run && count * sizeof(T) <= (1u << 20)) {
            void* r = slab_alloc(count * sizeof(T));
            if (r) return (T*)r;
        }
        size_t bytes = count * sizeof(T);
        if (bytes == 0) bytes = 16;
        if (!rezero_on_run) {
            void* r = table_alloc(bytes);
            if (r) {
                if (zero) cudaMemsetAsync(r, 0, bytes, engine->stream);
                return (T*)r;
            }
        }
        bool fresh = false;
        void* p = engine->dev_alloc(bytes, &fresh);
        if (!p) return nullptr;
        n_cuda_malloc += fresh ? 1 : 0;
        allocs.push_back(p);
        if (zero || rezero_on_run) cudaMemsetAsync(p, 0, bytes, engine->stream);
        if (rezero_on_run) zero_on_run.push_back({p, bytes});
        return (T*)p;
    }
    // Small uploads (instance tables, AudioParam timelines) have 4 MiB slabs of their own, are written into a host shadow of the slab and
    // sent with ONE copy per slab: flush_uploads(), at the end of a group's planning — before its first launch, on the stream the
    // launches go to.  (One cudaMemcpyAsync per table before: a plan with thousands of automated params issued thousands of them, each
    // a driver call that syncs the stream first because the source is pageable.)  Ranges are flushed once; tables of a group that is
    // still being planned by another worker may travel with this group's flush — they are complete (written under `mu`) and their
    // group flushes whatever it adds later.
    char* uslab = nullptr;
    std::vector<char> uslab_host;
    size_t uslab_used = 0, uslab_cap = 0, uslab_flushed = 0;
    void flush_uploads() {
        std::lock_guard<std::recursive_mutex> lk(mu);
        if (uslab && uslab_used > uslab_flushed) {
            cudaMemcpyAsync(uslab + uslab_flushed, uslab_host.data() + uslab_flushed, uslab_used - uslab_flushed, cudaMemcpyHostToDevice, engine->stream);
            uslab_flushed = uslab_used;
        }
    }
    void* upload_alloc(size_t bytes) {  // (under `mu`)
        bytes = (bytes + 255) / 256 * 256;
        if (bytes > (512u << 10)) return nullptr;
        if (!uslab || uslab_used + bytes > uslab_cap) {
            flush_uploads();  // what is left of the slab that is full
            bool fresh = false;
            void* p = engine->dev_alloc(4u << 20, &fresh);
            if (!p) return nullptr;
            n_cuda_malloc += fresh ? 1 : 0;
            allocs.push_back(p);
            uslab = (char*)p;
            uslab_cap = 4u << 20;
            uslab_used = uslab_flushed = 0;
            if (uslab_host.size() != uslab_cap) uslab_host.assign(uslab_cap, 0);  // (one shadow: the copy above has left it when cudaMemcpyAsync returns)
        }
        void* r = uslab + uslab_used;
        uslab_used += bytes;
        return r;
    }
    template <typename T>
    T* dupload(const std::vector<T>& v) {
        std::lock_guard<std::recursive_mutex> lk(mu);
        const size_t bytes = v.size() * sizeof(T);
        if (bytes > 0)
            if (void* r = upload_alloc(bytes)) {
                std::memcpy(uslab_host.data() + ((char*)r - uslab), v.data(), bytes);
                return (T*)r;
            }
        return dupload_now(v);
    }
    // ... and the direct form, for large tables and for data a kernel launched by the planner itself reads (the response of a convolver)
    template <typename T>
    T* dupload_now(const std::vector<T>& v) {
        std::lock_guard<std::recursive_mutex> lk(mu);
        T* p = dalloc<T>(v.size());
        if (p && !v.empty()) cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, engine->stream);
        return p;
    }
};

namespace {

// ---- topological order: Graph::order_nodes / visit (src/render/graph.rs:331-487) ---------------------------
// node id -> outgoing edges (ids are dense: a vector with presence flags; iteration in id order like the std::map it replaces).  The lists
// are the graph's own (no copies); a cycle breaker's list is replaced by the empty one.
struct EdgeTable {
    std::vector<const std::vector<Edge>*> v;
    std::vector<char> has;
    static const std::vector<Edge>& none() {
        static const std::vector<Edge> e;
        return e;
    }
    void reset(uint32_t max_id) {
        v.assign((size_t)max_id + 1, &none());
        has.assign((size_t)max_id + 1, 0);
    }
    void set(uint32_t id, const std::vector<Edge>* edges) {
        if (id >= v.size()) {
            v.resize((size_t)id + 1, &none());
            has.resize((size_t)id + 1, 0);
        }
        has[id] = 1;
        v[id] = edges;
    }
    void clear(uint32_t id) { set(id, &none()); }
    bool count(uint32_t id) const { return id < v.size() && has[id]; }
    const std::vector<Edge>* find(uint32_t id) const { return count(id) ? v[id] : nullptr; }
    const std::vector<Edge>& at(uint32_t id) const {
        if (!count(id)) throw std::out_of_range("orderer: unknown node id");
        return *v[id];
    }
};

struct Orderer {
    wae_graph* g;
    EdgeTable edges;  // working copy: cycle breakers clear a DelayWriter's edges
    std::vector<uint32_t> ordered, marked, marked_temp, in_cycle, cycle_breakers, broken;
    static bool contains(const std::vector<uint32_t>& v, uint32_t x) { return std::find(v.begin(), v.end(), x) != v.end(); }
    // returns true when a cycle breaker was applied (the ordering is then restarted), graph.rs:331-403.  Same visiting order
    // as the reference; membership tests use flag vectors indexed by node id instead of its linear `contains` (O(n^2) on 10^4-node graphs).
    struct IdSet {
        std::vector<char> f;
        void reset(size_t n) { f.assign(n, 0); }
        bool count(uint32_t id) const { return id < f.size() && f[id]; }
        bool insert(uint32_t id) {  // true: newly inserted
            if (id >= f.size()) f.resize((size_t)id + 1, 0);
            const bool fresh = !f[id];
            f[id] = 1;
            return fresh;
        }
        void erase(uint32_t id) { if (id < f.size()) f[id] = 0; }
    };
    IdSet marked_set, temp_set;
    bool visit(uint32_t id) {
        if (temp_set.count(id)) {
            auto it = std::find(marked_temp.begin(), marked_temp.end(), id);
            for (auto jt = it; jt != marked_temp.end(); ++jt)
                if (g->nodes.at(*jt).cycle_breaker) {
                    cycle_breakers.push_back(*jt);
                    return true;
                }
            in_cycle.insert(in_cycle.end(), it, marked_temp.end());  // no DelayNode in the cycle: its nodes are muted
            return false;
        }
        if (!marked_set.insert(id)) return false;
        marked_temp.push_back(id);
        temp_set.insert(id);
        const std::vector<Edge>& out = edges.at(id);
        for (size_t i = 0; i < out.size(); i++)
            if (edges.count(out[i].other_id) && visit(out[i].other_id)) return true;  // (every node of the graph has an entry)
        ordered.push_back(id);
        // `id` is the innermost node still being visited: it is the last entry of the stack unless an unbroken cycle was
        // recorded below it, in which case the reference's `retain` removes it wherever it is
        if (!marked_temp.empty() && marked_temp.back() == id) marked_temp.pop_back();
        else marked_temp.erase(std::remove(marked_temp.begin(), marked_temp.end(), id), marked_temp.end());
        temp_set.erase(id);
        return false;
    }
    void run() {  // graph.rs:418-487
        edges.reset(g->nodes.empty() ? 0 : g->nodes.max_id());
        for (auto& kv : g->nodes) edges.set(kv.first, &kv.second.outgoing);
        ordered.reserve(g->nodes.size());
        for (;;) {
            ordered.clear(); marked.clear(); marked_temp.clear(); in_cycle.clear(); cycle_breakers.clear();
            marked_set.reset(edges.v.size()); temp_set.reset(edges.v.size());
            bool applied = false;
            for (auto& kv : g->nodes) {
                applied = visit(kv.first);
                if (applied) break;
            }
            if (!applied) break;
            for (uint32_t id : cycle_breakers) {
                edges.clear(id);
                if (!contains(broken, id)) broken.push_back(id);
            }
        }
        if (!in_cycle.empty()) {
            std::unordered_set<uint32_t> muted(in_cycle.begin(), in_cycle.end());
            ordered.erase(std::remove_if(ordered.begin(), ordered.end(), [&](uint32_t o) { return muted.count(o) != 0; }), ordered.end());
        }
        std::reverse(ordered.begin(), ordered.end());
    }
};

struct PortRef {
    uint32_t node;
    int port;
};

// What the planner can say about a buffer's layout over the render (see BufRef::meta): the range of its channel count over all quanta
// (a silent quantum has one channel, quantum.rs:512-517, unless it sits in a port with an explicit count), the range over the quanta
// that are not silent, and whether it can be silent at all.  A layout that is provably constant needs no track and keeps every
// kernel on its static path; sources that run from frame 0 to the end of the render (every BASELINE config) are.
struct Lay {
    uint8_t lo = 1, hi = 1, nlo = 1, nhi = 1;
    bool may_silent = false;
    bool dyn() const { return lo != hi || may_silent; }
    static Lay fixed(int ch) { return Lay{(uint8_t)ch, (uint8_t)ch, (uint8_t)ch, (uint8_t)ch, false}; }
    static Lay gated(int ch) { return Lay{1, (uint8_t)ch, (uint8_t)ch, (uint8_t)ch, true}; }  // `ch` channels or silent
};

struct PNode {
    Node* n = nullptr;
    int level = 0;
    std::vector<std::vector<PortRef>> in_edges;  // per input port, reference summation order
    std::vector<int> in_ch;
    std::vector<BufRef> in_buf;
    std::vector<Lay> in_lay;
    std::vector<int> out_ch;
    std::vector<BufRef> out_buf;
    std::vector<Lay> out_lay;  // empty: constant (out_ch channels, never silent)
    bool wrote_dest = false;   // out_buf[0] IS the graph's rendered PCM (a convolver that is the destination's only input)
    Lay lay_out(int port) const { return port < (int)out_lay.size() ? out_lay[port] : Lay::fixed(out_ch[port]); }
};

// node id -> PNode; ids are handed out densely (wae_graph::next_id), so this is a vector, not a tree (the planner looks nodes up
// several times per edge).  One table per Planner, reset from graph to graph: the per-node vectors keep their capacity, so the graphs
// of a batch after the first are planned without allocating them again (seven vectors per node).
struct NodeTable {
    std::vector<PNode> v;
    std::vector<char> has;
    void reset(uint32_t max_id) {
        if (v.size() < (size_t)max_id + 1) v.resize((size_t)max_id + 1);
        has.assign(v.size(), 0);
    }
    PNode& put(uint32_t id, Node* n) {
        if (id >= v.size()) {
            v.resize((size_t)id + 1);
            has.resize((size_t)id + 1, 0);
        }
        PNode& p = v[id];
        p.n = n;
        p.level = 0;
        p.in_edges.resize((size_t)n->n_inputs);
        for (auto& port : p.in_edges) port.clear();
        p.in_ch.clear();
        p.in_buf.clear();
        p.in_lay.clear();
        p.out_ch.clear();
        p.out_buf.clear();
        p.out_lay.clear();
        p.wrote_dest = false;
        has[id] = 1;
        return p;
    }
    PNode* find(uint32_t id) { return id < v.size() && has[id] ? &v[id] : nullptr; }
    bool count(uint32_t id) const { return id < v.size() && has[id]; }
    PNode& at(uint32_t id) {
        if (!(id < v.size() && has[id])) throw std::out_of_range("planner: unknown node id");
        return v[id];
    }
};

struct Planner {
    wae_batch* b;
    wae_engine* eng;
    std::map<std::pair<int, int>, StageBuild> builds;  // (level, kind * 64 + variant)
    std::string error;
    int error_code = 0;
    uint64_t algorithmic_bytes = 0;
    // IR spectra cache: content hash -> device spectra (shared by the planners of all groups of a batch, guarded by b->mu)
    struct IrSpectra {
        float2* h;
        int S;
        int channels;
    };
    std::unordered_map<uint64_t, IrSpectra> own_ir_cache;
    std::unordered_map<uint64_t, IrSpectra>* ir_cache = &own_ir_cache;
    std::map<int, std::pair<const float2*, const float2*>> os_filters;  // over-sampled shaper: factor -> (up, down) filter bins

    bool has_feedback = false;            // some graph has a cycle broken by a DelayNode
    std::map<std::pair<uint32_t, uint32_t>, int>* delay_ch_hint = nullptr;  // (graph, reader id) -> channels of an in-cycle delay
    std::map<std::pair<uint32_t, uint32_t>, int> delay_ch_seen;
    struct DelayRing {
        float* ring;
        uint32_t ring_len;
        int ch;
        int64_t* mono_at;
        int32_t mono_len;
    };
    std::map<std::pair<uint32_t, uint32_t>, DelayRing> delay_rings;       // (graph, writer id)
    bool dry = false;                     // sizing pass: count arena floats per frame, touch no device memory
    int group_graphs = 1;                 // graphs of the group being planned (k_voice_sum: are there enough work items?)
    // 0 off (default: measured slower than k_chain + k_mix on north_star, profiles/README.md r2_q / r2_r), 1 when the launch is large enough,
    // 2 whenever the port has the shape (tests).  WAE_OPT_VOICE_SUM, else WAE_VOICE_SUM from the environment (read per plan).
    int vs_mode = -1;
    int voice_sum_mode() {
        if (eng->voice_sum >= 0) return eng->voice_sum;
        if (vs_mode < 0) {
            const char* e = getenv("WAE_VOICE_SUM");
            vs_mode = e ? std::max(0, std::min(2, atoi(e))) : 0;
        }
        return vs_mode;
    }
    uint64_t arena_floats_per_frame = 0;
    // source PCM slab of the group being planned (device pointer, pinned host mirror, cursor in floats)
    float* d_src = nullptr;
    std::vector<wae_batch::Group::SrcCopy>* src_copies = nullptr;
    size_t src_cursor = 0;
    struct PendingChain {
        ChainInst inst;
        hm::BiquadCoefs coefs[CHAIN_MAX_BIQUADS] = {};  // of inst.bq[k]: the scan constants (1.3 KB a set) are derived when the chain is emitted
        int ch = 1;
        int phase = 0;  // 0: before biquad A, 1: after A, 3: after B, 5: after the shaper (canonical chain order)
        int cls = 0;    // scheduling class of the node that opened the chain (see stage())
        Lay lay;        // layout of the chain's output over time (the kernel writes the track when it is not constant)
    };

    // Node state is allocated through a key (graph, node, n-th allocation of that node, salt): the plans of consecutive
    // render segments (suspend_sync) find the state of a node that lives on, new nodes get fresh (zeroed) state.
    uint32_t key_graph = 0, key_node = 0, key_seq = 0;
    uint64_t key_salt = 0;
    template <typename T>
    T* alloc(size_t count, bool zero = false, bool rezero_on_run = false) {
        if (dry) return reinterpret_cast<T*>(uintptr_t(256));
        if (seg_start == 0 && seg_end >= b->lq) return b->dalloc<T>(count, zero, rezero_on_run);  // no suspend point: no later plan looks it up
        const wae_batch::StateKey key{key_graph, key_node, key_seq++, key_salt};
        const size_t bytes = count * sizeof(T);
        std::lock_guard<std::recursive_mutex> lk(b->mu);
        auto it = b->state_map.find(key);
        if (it != b->state_map.end() && it->second.second == bytes) return (T*)it->second.first;
        T* p = b->dalloc<T>(count, zero, rezero_on_run);
        if (p) b->state_map[key] = {(void*)p, bytes};
        return p;
    }
    // arena buffers are chunk-local scratch: every segment's plan draws from the same pool
    std::map<int, std::vector<float*>> arena_pool;
    std::map<int, size_t> arena_used;
    std::map<std::pair<uint32_t, uint32_t>, size_t> src_offsets;  // (graph, buffer source node) -> offset in the group's PCM slab
    std::unordered_map<const PcmBuffer*, size_t> buf_offsets;       // one copy per AudioBuffer in the slab, whatever number of nodes play it
    // render-side view of every AudioParam: the event queue it (re)started with at `init_frame`.  When a suspend callback
    // pushed more events, the state machine is replayed on the host up to the suspend frame and the new events are folded
    // into what is left of the queue — handle_incoming_event against the live state, like the reference's render thread.
    struct ParamRecord {
        size_t n_source_events = 0;  // arrival-order events of the Param already folded in
        ParamTimeline tl;
        int64_t init_frame = 0;
    };
    std::unordered_map<uint64_t, ParamRecord> param_records;  // (graph << 32 | param id); element addresses survive rehashing
    const ParamTimeline* param_timeline(uint32_t gi, uint32_t pid, const Param& prm, float sample_rate) {
        const uint64_t rec_key = (uint64_t)gi << 32 | pid;
        auto it = param_records.find(rec_key);
        if (it == param_records.end()) {
            ParamRecord r;
            r.tl = build_param_timeline(prm);
            r.n_source_events = prm.events.size();
            r.init_frame = seg_start;
            return &param_records.emplace(rec_key, std::move(r)).first->second.tl;
        }
        ParamRecord& r = it->second;
        if (r.n_source_events == prm.events.size() || !r.tl.error.empty()) return &r.tl;
        // replay compute_buffer from the record's start to this segment's start
        ParamInst host{};
        host.events = r.tl.events.data();
        host.curves = r.tl.curves.data();
        host.n_events = (int32_t)r.tl.events.size();
        host.a_rate = prm.a_rate ? 1 : 0;
        host.sample_rate = sample_rate;
        ParamState st{};
        st.intrinsic = r.tl.intrinsic;
        st.has_last = r.tl.has_last ? 1 : 0;
        st.last = r.tl.last;
        st.inited = 1;
        float buf[128];
        for (int64_t f = r.init_frame; f < seg_start; f += 128) param_compute_buffer(host, st, (double)f / (double)sample_rate, buf);
        ParamTimeline next;
        next.curves = r.tl.curves;
        for (int i = st.head; i < host.n_events; i++) next.events.push_back(i == st.head && st.override_valid ? st.override_ev : r.tl.events[i]);
        next.intrinsic = st.intrinsic;
        next.has_last = st.has_last != 0;
        next.last = st.last;
        fold_param_events(next, prm.events.data() + r.n_source_events, prm.events.size() - r.n_source_events);
        r.tl = std::move(next);
        r.n_source_events = prm.events.size();
        r.init_frame = seg_start;
        return &r.tl;
    }
    int64_t seg_start = 0, seg_end = 0;
    void begin_segment(int64_t f0, int64_t f1) {
        seg_start = f0;
        seg_end = f1;
        builds.clear();
        arena_used.clear();
        arena_floats_per_frame = 0;
    }
    template <typename T>
    T* upload(const std::vector<T>& v) {
        if (dry) return reinterpret_cast<T*>(uintptr_t(256));
        return b->dupload(v);
    }

    bool bail(int code, const std::string& msg) {
        if (!error_code) {
            error_code = code;
            error = msg;
        }
        return false;
    }

    // Scheduling class of a stage.  Graphs without DelayNode feedback: 0 (whole chunks).  Graphs with feedback: 0 = strictly
    // upstream of every cycle (whole chunks, rendered first: sources, a reverb feeding an echo loop), 1 = on a path from a
    // cycle to a cycle-breaking DelayWriter, i.e. inside a feedback cycle or between two of them (replayed quantum by quantum
    // inside the chunk, like the reference's render loop), 2 = the rest: downstream of the cycles only (whole chunks again,
    // e.g. a reverb after an echo loop).
    int cur_cls = 0;  // class of the node being planned
    StageBuild& stage(int level, int kind, int variant = 0) {
        const int cls = cur_cls;
        StageBuild& s = builds[{cls * 1000000 + level, kind * 64 + variant}];
        s.cls = cls;
        s.level = level;
        s.kind = kind;
        s.variant = variant;
        return s;
    }

    // `with_meta`: the buffer's layout is not provably constant: it carries a per-quantum layout track (BufRef::meta), one row per
    // static channel, stored behind the PCM
    BufRef arena_buf(int ch, bool with_meta = false) {
        arena_floats_per_frame += (uint64_t)ch;
        const uint32_t mstride = (uint32_t)((b->chunk / 128 + 16) / 16 * 16);
        BufRef r{reinterpret_cast<float*>(uintptr_t(256)), (uint32_t)b->chunk, 0, nullptr, 0, 0};
        if (!dry) {
            std::vector<float*>& pool = arena_pool[ch];
            size_t& used = arena_used[ch];
            if (used < pool.size()) {
                r.p = pool[used++];
            } else {
                const size_t floats = (size_t)ch * (size_t)b->chunk;
                float* p = b->dalloc<float>(floats + ((size_t)ch * mstride + 3) / 4);
                b->arena_bytes += floats * 4;
                if (p) {
                    pool.push_back(p);
                    used++;
                }
                r.p = p;
            }
        }
        if (with_meta && r.p) {
            r.meta = reinterpret_cast<uint8_t*>(r.p + (size_t)ch * (size_t)b->chunk);
            r.meta_stride = mstride;
        }
        return r;
    }
    // layout track of a node output from its input's (k_meta)
    void meta_stage(int L, int mode, const BufRef& in, int in_ch, const BufRef& out, int out_ch, int count = 0, int aux = 0) {
        MetaInst m{};
        m.in = in;
        m.out = out;
        m.mode = mode;
        m.in_ch = in_ch;
        m.out_ch = out_ch;
        m.count = count;
        m.aux = aux;
        stage(L, S_META).meta.push_back(m);
    }

    NodeTable node_table;  // of the graph being planned (reused from graph to graph)
    NodeTable* cur_pn = nullptr;
    struct PRef {
        bool dyn = false;  // automated / audio-rate driven: one value per frame in `track`
        float v = 0.f;
        BufRef track{nullptr, 0, 0};
    };
    PRef param_ref(wae_graph* g, uint32_t pid);
    bool plan_graph(wae_graph* g, uint32_t gi);
    // ir_override: the response of a STATIC HRTF panner (blended, gain folded in): no normalisation, no trimming of small trailing taps;
    // a two-channel input is mixed down to mono by the forward transform's loads (ConvInput::in_channel = -1)
    bool plan_convolver(wae_graph* g, PNode& pn, int level, const BufRef* dest = nullptr, int64_t dest_limit = -1, const PcmBuffer* ir_override = nullptr);
};

static uint64_t fnv1a(const void* data, size_t bytes, uint64_t h = 1469598103934665603ull) {
    const uint8_t* p = (const uint8_t*)data;
    // 8 bytes at a time is enough for a cache key
    size_t n8 = bytes / 8;
    const uint64_t* q = (const uint64_t*)p;
    for (size_t i = 0; i < n8; i++) {
        h ^= q[i];
        h *= 1099511628211ull;
    }
    for (size_t i = n8 * 8; i < bytes; i++) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}

// computedNumberOfChannels for one input port (src/render/quantum.rs:543-547), static channel counts
// Developer check of planner refactorings (WAE_PLAN_DIGEST=1, wae_batch_plan only): a hash over every instance record the sizing pass
// builds (records are value-initialised, device pointers are the dry pass's placeholders), printed per group — equal digests before and
// after a change of the planner's data structures mean the same tables would be uploaded.
static bool plan_digest_wanted() {
    static const bool on = [] { const char* e = getenv("WAE_PLAN_DIGEST"); return e && atoi(e) != 0; }();
    return on;
}
template <typename T>
static uint64_t digest_vec(const std::vector<T>& v, uint64_t h) {
    const uint64_t n = v.size();
    h = fnv1a(&n, sizeof n, h);
    if (v.empty()) return h;
    static const bool dump = [] { const char* e = getenv("WAE_PLAN_DIGEST"); return e && atoi(e) >= 2; }();
    if (dump) {  // which 8-byte word of which record type differs between two runs
        std::fprintf(stderr, "  [%s] n %zu size %zu:", __PRETTY_FUNCTION__, v.size(), sizeof(T));
        const uint64_t* q = (const uint64_t*)v.data();
        for (size_t i = 0; i < sizeof(T) / 8 && i < 400; i++) std::fprintf(stderr, " %llx", (unsigned long long)q[i]);
        std::fprintf(stderr, "\n");
    }
    return fnv1a(v.data(), v.size() * sizeof(T), h);
}
static uint64_t digest_builds(const std::map<std::pair<int, int>, StageBuild>& builds, uint64_t h) {
    for (auto& kv : builds) {
        const StageBuild& s = kv.second;
        const int key[6] = {kv.first.first, kv.first.second, s.cls, s.level, s.kind * 64 + s.variant, s.max_ch};
        h = fnv1a(key, sizeof key, h);
        h = digest_vec(s.osc, h); h = digest_vec(s.cst, h); h = digest_vec(s.absn, h); h = digest_vec(s.biquad, h); h = digest_vec(s.chain, h);
        h = digest_vec(s.param, h); h = digest_vec(s.osc_ar, h); h = digest_vec(s.biquad_ar, h); h = digest_vec(s.absn_slow, h);
        h = digest_vec(s.scan_coef, h); h = digest_vec(s.iir, h); h = digest_vec(s.gain, h); h = digest_vec(s.shaper, h); h = digest_vec(s.span, h);
        h = digest_vec(s.span_gains, h); h = digest_vec(s.pan, h); h = digest_vec(s.hrtf, h); h = digest_vec(s.hrtf_sel, h); h = digest_vec(s.pan_dyn, h);
        h = digest_vec(s.absn_serial, h); h = digest_vec(s.shaper_os, h); h = digest_vec(s.route, h); h = digest_vec(s.delay, h); h = digest_vec(s.comp, h);
        h = digest_vec(s.analyser, h); h = digest_vec(s.mix, h); h = digest_vec(s.mix_edges, h); h = digest_vec(s.mix_dyn, h); h = digest_vec(s.meta, h);
        h = digest_vec(s.conv_in, h); h = digest_vec(s.conv_path, h); h = digest_vec(s.vgroups, h);
    }
    return h;
}

// Split planning (a group of few, large graphs planned by several workers, each with its own Planner over a contiguous run of the
// group's graphs): the runs' stage builds are appended to one another in graph order, which is the order a single planner would have
// produced.  Records that index a sibling table of their stage are rebased: mix instances -> mix edges, chain biquads -> scan constants,
// voice groups -> chain records, convolver paths -> the conv-input table of the forward-transform stage of their level.
using Builds = std::map<std::pair<int, int>, StageBuild>;
template <typename T>
static void append_vec(std::vector<T>& d, std::vector<T>& s) {
    if (d.empty()) d = std::move(s);
    else d.insert(d.end(), s.begin(), s.end());
}
static void merge_builds(Builds& dst, Builds& src) {
    struct Base {
        size_t mix_edges = 0, scan = 0, chain = 0, conv_in = 0;
    };
    std::map<std::pair<int, int>, Base> base;  // table sizes of `dst` before anything of `src` is appended
    for (auto& kv : src) {
        auto it = dst.find(kv.first);
        if (it != dst.end()) base[kv.first] = Base{it->second.mix_edges.size(), it->second.n_scan_coef, it->second.chain.size(), it->second.conv_in.size()};
        else base[kv.first] = Base{};
    }
    for (auto& kv : src) {
        StageBuild& s = kv.second;
        const Base bs = base[kv.first];
        for (auto& m : s.mix) m.edge_offset += (uint32_t)bs.mix_edges;
        for (auto& m : s.mix_dyn) m.edge_offset += (uint32_t)bs.mix_edges;
        for (auto& c : s.chain)
            for (int k = 0; k < c.n_biquad; k++) c.bq[k].coef += (int32_t)bs.scan;
        for (auto& v : s.vgroups) v.first += (int32_t)bs.chain;
        if (s.kind == S_CONV_MAC || s.kind == S_CONV_MAC_ACC) {
            const std::pair<int, int> fft_key{kv.first.first, S_CONV_FFT * 64};
            size_t off = 0;
            auto bi = base.find(fft_key);
            if (bi != base.end()) off = bi->second.conv_in;
            else if (dst.count(fft_key)) off = dst[fft_key].conv_in.size();
            for (auto& cp : s.conv_path) cp.input += (int32_t)off;
        }
        auto it = dst.find(kv.first);
        if (it == dst.end()) {
            dst.emplace(kv.first, std::move(s));
            continue;
        }
        StageBuild& d = it->second;
        append_vec(d.osc, s.osc); append_vec(d.cst, s.cst); append_vec(d.absn, s.absn); append_vec(d.biquad, s.biquad); append_vec(d.chain, s.chain);
        append_vec(d.param, s.param); append_vec(d.osc_ar, s.osc_ar); append_vec(d.biquad_ar, s.biquad_ar); append_vec(d.absn_slow, s.absn_slow);
        append_vec(d.scan_coef, s.scan_coef); append_vec(d.iir, s.iir); append_vec(d.gain, s.gain); append_vec(d.shaper, s.shaper); append_vec(d.span, s.span);
        append_vec(d.span_gains, s.span_gains); append_vec(d.pan, s.pan); append_vec(d.hrtf, s.hrtf); append_vec(d.hrtf_sel, s.hrtf_sel);
        append_vec(d.pan_dyn, s.pan_dyn); append_vec(d.absn_serial, s.absn_serial); append_vec(d.shaper_os, s.shaper_os); append_vec(d.route, s.route);
        append_vec(d.delay, s.delay); append_vec(d.comp, s.comp); append_vec(d.analyser, s.analyser); append_vec(d.mix, s.mix);
        append_vec(d.mix_edges, s.mix_edges); append_vec(d.mix_dyn, s.mix_dyn); append_vec(d.meta, s.meta); append_vec(d.conv_in, s.conv_in);
        append_vec(d.conv_path, s.conv_path); append_vec(d.vgroups, s.vgroups);
        d.n_scan_coef += s.n_scan_coef;
        d.max_ch = std::max(d.max_ch, s.max_ch);
    }
}

static int computed_channels(const ChannelCfg& cfg, int max_in) {
    switch (cfg.mode) {
        case WAE_COUNT_MODE_MAX: return max_in;
        case WAE_COUNT_MODE_EXPLICIT: return cfg.count;
        default: return std::min(max_in, cfg.count);
    }
}

static uint32_t next_pow2(uint64_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

// constants of the time-parallel biquad recurrence (see ScanCoef in wae_kernels.h), f64 on the host
static ScanCoef make_scan_coef(const hm::BiquadCoefs& c) {
    ScanCoef sc{};
    struct M2 {
        double a, b, c, d;
    };
    auto mul = [](const M2& x, const M2& y) { return M2{x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d}; };
    const M2 M{-c.a1, -c.a2, 1., 0.};
    M2 r{1., 0., 0., 1.};
    for (int j = 0; j < WAE_CHAIN_K; j++) r = mul(M, r);
    const M2 A = r;  // M^K: one thread of k_chain
    M2 pw = A;
    for (int d = 0; d < 5; d++) {
        sc.Pshfl[d][0] = pw.a; sc.Pshfl[d][1] = pw.b; sc.Pshfl[d][2] = pw.c; sc.Pshfl[d][3] = pw.d;
        pw = mul(pw, pw);
    }
    sc.Pwarp[0] = pw.a; sc.Pwarp[1] = pw.b; sc.Pwarp[2] = pw.c; sc.Pwarp[3] = pw.d;  // A^32
    M2 pl = A;
    for (int l = 0; l < 32; l++) {
        sc.Plane[l][0] = pl.a; sc.Plane[l][1] = pl.b; sc.Plane[l][2] = pl.c; sc.Plane[l][3] = pl.d;
        pl = mul(A, pl);
    }
    // one frame with zero input: (x1, x2, y1, y2) -> (0, x1, b1 x1 + b2 x2 - a1 y1 - a2 y2, y1); G^L by squaring (L = 2^15 frames)
    static_assert(WAE_CHAIN_PRE_TILES * WAE_CHAIN_K * 128 == 1 << 15, "GL below is G^(2^15)");
    double g[16] = {0., 0., 0., 0., 1., 0., 0., 0., c.b1, c.b2, -c.a1, -c.a2, 0., 0., 1., 0.}, h[16];
    for (int sq = 0; sq < 15; sq++) {
        for (int r = 0; r < 4; r++)
            for (int cc = 0; cc < 4; cc++) {
                double a = 0.;
                for (int k = 0; k < 4; k++) a += g[4 * r + k] * g[4 * k + cc];
                h[4 * r + cc] = a;
            }
        std::memcpy(g, h, sizeof g);
    }
    std::memcpy(sc.GL, g, sizeof g);
    return sc;
}

bool Planner::plan_convolver(wae_graph* g, PNode& pn, int level, const BufRef* dest, int64_t dest_limit, const PcmBuffer* ir_override) {
    Node& n = *pn.n;
    int in_ch = pn.in_ch[0];
    const bool mono_mix = ir_override && in_ch == 2;
    if (mono_mix) in_ch = 1;
    if ((n.buffer || ir_override) && cur_cls == 1)  // (the class is a property of the graph: the sizing pass already knows it)
        return bail(WAE_UNSUPPORTED, "a ConvolverNode inside a DelayNode feedback cycle is not lowered to the GPU (before or after the cycle it is)");
    const Lay in_lay = pn.in_lay.empty() ? Lay::fixed(in_ch) : pn.in_lay[0];
    if (!n.buffer && !ir_override) {  // no buffer: pass-through (convolver.rs:368-375)
        pn.out_ch = {in_ch};
        pn.out_buf = {pn.in_buf[0]};
        pn.out_lay = {in_lay};
        return true;
    }
    const PcmBuffer& ir = ir_override ? *ir_override : *n.buffer;
    int ir_ch = (int)ir.channels.size();
    size_t ir_len = ir.length();
    // normalize_buffer, src/node/convolver.rs:16-53 (f32, channel by channel)
    float scale = 1.f;
    if (n.normalize && !ir_override) {
        float power = 0.f;
        for (auto& c : ir.channels) {
            float s = 0.f;
            for (float v : c) s += v * v;
            power += s;
        }
        power = std::sqrt(power / (float)((size_t)ir_ch * ir_len));
        if (!std::isfinite(power) || power < 0.000125f) power = 0.000125f;
        scale = 1.f / power;
        scale *= 0.00125f;
        scale *= 44100.f / ir.sample_rate;
        if (ir_ch == 4) scale *= 0.5f;
    }
    // convolvers: one per IR channel, a mono IR is duplicated (convolver.rs:289-293)
    int n_conv = std::max(ir_ch, 2);
    // trailing samples below 1e-6 are ignored by fft-convolver's init
    std::vector<std::vector<float>> scaled(ir_ch);
    size_t trimmed_len = 0;
    for (int c = 0; c < ir_ch; c++) {
        scaled[c].resize(ir_len);
        for (size_t i = 0; i < ir_len; i++) scaled[c][i] = ir.channels[c][i] * scale;
    }
    // per-channel trimmed length (each FFTConvolver trims its own IR); use per channel S
    std::vector<int> S(ir_ch);
    for (int c = 0; c < ir_ch; c++) {
        size_t m = ir_len;
        while (!ir_override && m > 0 && std::fabs(scaled[c][m - 1]) < 0.000001f) m--;
        while (ir_override && m > 0 && scaled[c][m - 1] == 0.f) m--;  // (exact zeros only)
        S[c] = (int)((m + WAE_CONV_BLOCK - 1) / WAE_CONV_BLOCK);
        trimmed_len = std::max(trimmed_len, m);
        // zero the ignored tail so that a shared segment count reproduces the per-convolver trimming
        for (size_t i = m; i < ir_len; i++) scaled[c][i] = 0.f;
    }
    int Smax = *std::max_element(S.begin(), S.end());
    pn.out_ch = {ir_ch == 1 && in_ch == 1 ? 1 : 2};
    if (in_lay.dyn() && ir_ch == 1 && in_lay.hi >= 2)
        return bail(WAE_UNSUPPORTED, "a ConvolverNode with a mono response whose input changes between one and two channels is not lowered to the GPU "
                                     "(the reference stops feeding its second convolver whenever the input is mono or silent, convolver.rs:378-400)");
    // silent once the input has been silent for the length of the response (convolver.rs:357-366); channels from the routing table (:378-487)
    const bool conv_dyn = in_lay.dyn();
    // the destination's only input, same channel count, constant layout: the inverse transforms write the rendered PCM themselves
    const bool direct = dest && !conv_dyn && Smax > 0 && pn.out_ch[0] == (int)b->channels;
    pn.out_buf = {direct ? *dest : arena_buf(pn.out_ch[0], conv_dyn && Smax > 0)};
    pn.wrote_dest = direct;
    if (conv_dyn && Smax > 0) {
        const int oc = pn.out_ch[0];
        pn.out_lay = {Lay{(uint8_t)(in_lay.may_silent ? 1 : oc), (uint8_t)oc, (uint8_t)oc, (uint8_t)oc, in_lay.may_silent}};
        MetaInst m{};
        m.in = pn.in_buf[0];
        m.out = pn.out_buf[0];
        m.mode = META_CONV;
        m.in_ch = in_ch;
        m.out_ch = oc;
        m.aux = ir_ch;
        m.tail_len = (int64_t)ir_len;
        m.state = alloc<int64_t>(1, true, true);
        if (!m.state) return bail(WAE_OUT_OF_MEMORY, "out of device memory (convolver tail counter)");
        stage(level, S_META).meta.push_back(m);
    }
    if (Smax == 0) {  // all-zero IR: output zeros -> a mix with no edges
        StageBuild& ms = stage(level, S_MIX);
        ms.mix.push_back(MixInst{pn.out_buf[0], pn.out_ch[0], 0, 0, (uint32_t)ms.mix_edges.size(), -1});
        return true;
    }
    // IR spectra (deduplicated across the batch by content)
    uint64_t key = fnv1a(&scale, sizeof(scale));
    for (int c = 0; c < ir_ch; c++) key = fnv1a(ir.channels[c].data(), ir_len * sizeof(float), key);
    key = fnv1a(&ir_len, sizeof(ir_len), key);
    IrSpectra spec;
    std::unique_lock<std::recursive_mutex> ir_lock(b->mu);
    auto it = ir_cache->find(key);
    if (it != ir_cache->end()) {
        spec = it->second;
    } else {
        std::vector<float> flat((size_t)ir_ch * ir_len);
        for (int c = 0; c < ir_ch; c++) std::memcpy(flat.data() + (size_t)c * ir_len, scaled[c].data(), ir_len * sizeof(float));
        float* d_ir = dry ? upload(flat) : b->dupload_now(flat);  // (read by launch_conv_ir_fft below: not through the deferred upload slabs)
        if (!dry) cudaStreamSynchronize(eng->stream);  // `flat` is about to go out of scope
        spec.S = Smax;
        spec.channels = ir_ch;
        spec.h = alloc<float2>((size_t)ir_ch * (Smax + WAE_CONV_H_PAD) * WAE_CONV_SPEC, true);  // (zeroed: the padding partitions of every channel)
        if (!d_ir || !spec.h) return bail(WAE_OUT_OF_MEMORY, "out of device memory (IR spectra)");
        if (!dry) launch_conv_ir_fft(d_ir, (int64_t)ir_len, (int64_t)ir_len, spec.h, Smax, ir_ch, eng->stream);
        b->asset_bytes += (size_t)ir_ch * (Smax + WAE_CONV_H_PAD) * WAE_CONV_SPEC * 8;
        (*ir_cache)[key] = spec;
    }
    ir_lock.unlock();
    // inputs: one spectra ring per input channel
    StageBuild& fs = stage(level, S_CONV_FFT);
    int blocks_per_chunk = (int)((b->chunk + WAE_CONV_BLOCK - 1) / WAE_CONV_BLOCK);
    int ring_blocks = Smax + blocks_per_chunk;
    int in_base = (int)fs.conv_in.size();
    for (int c = 0; c < in_ch; c++) {
        ConvInput ci;
        ci.in = pn.in_buf[0];
        ci.in_channel = mono_mix ? -1 : c;
        ci.prev = alloc<float>(WAE_CONV_BLOCK, true, true);
        ci.xring = alloc<float2>((size_t)ring_blocks * WAE_CONV_SPEC);
        ci.xring_blocks = ring_blocks;
        if (!ci.prev || !ci.xring) return bail(WAE_OUT_OF_MEMORY, "out of device memory (convolver input spectra)");
        b->arena_bytes += (size_t)ring_blocks * WAE_CONV_SPEC * 8;
        fs.conv_in.push_back(ci);
    }
    // paths: channel routing table of convolver.rs:378-487
    struct R {
        int in, ir, out, acc;
    };
    std::vector<R> routes;
    if (in_ch == 1 && ir_ch == 1) routes = {{0, 0, 0, 0}};
    else if (in_ch == 1 && ir_ch == 2) routes = {{0, 0, 0, 0}, {0, 1, 1, 0}};
    else if (in_ch == 2 && ir_ch == 1) routes = {{0, 0, 0, 0}, {1, 0, 1, 0}};
    else if (in_ch == 2 && ir_ch == 2) routes = {{0, 0, 0, 0}, {1, 1, 1, 0}};
    else if (in_ch == 2 && ir_ch == 4) routes = {{0, 0, 0, 0}, {0, 1, 1, 0}, {1, 2, 0, 1}, {1, 3, 1, 1}};
    else if (in_ch == 1 && ir_ch == 4) routes = {{0, 0, 0, 0}, {0, 1, 1, 0}, {0, 2, 0, 1}, {0, 3, 1, 1}};
    else return bail(WAE_UNSUPPORTED, "unsupported convolver channel routing");
    (void)n_conv;
    for (auto& r : routes) {
        ConvPath p;
        p.out = pn.out_buf[0];
        p.h = spec.h + ((size_t)r.ir * (Smax + WAE_CONV_H_PAD) + WAE_CONV_H_PAD_LO) * WAE_CONV_SPEC;
        p.input = in_base + r.in;
        p.S = Smax;
        p.out_channel = r.out;
        p.accumulate = r.acc;
        p.limit = direct ? dest_limit : -1;
        p.y = nullptr;
        if (Smax > 1) {  // (one partition: k_conv_ifft forms the product itself, there are no output spectra)
            p.y = alloc<float2>((size_t)blocks_per_chunk * WAE_CONV_SPEC);
            if (!p.y) return bail(WAE_OUT_OF_MEMORY, "out of device memory (convolver output spectra)");
            b->arena_bytes += (size_t)blocks_per_chunk * WAE_CONV_SPEC * 8;
        }
        stage(level, r.acc ? S_CONV_MAC_ACC : S_CONV_MAC).conv_path.push_back(p);
    }
    // S*1025*8 B of input-history spectra per convolver-block of 1024 frames
    // (the reference's 1024-frame partitioning defines the algorithmic figure, whatever block size the kernels use)
    if (!ir_override) algorithmic_bytes += (uint64_t)routes.size() * (uint64_t)((trimmed_len + 1023) / 1024) * 1025ull * 8ull * (uint64_t)((b->lq + 1023) / 1024);
    return true;
}

Planner::PRef Planner::param_ref(wae_graph* g, uint32_t pid) {
    PRef r;
    PNode* it = cur_pn->find(pid);
    r.v = (it ? *it->n : g->nodes.at(pid)).param.constant_value();
    if (it && !it->out_buf.empty()) {
        r.dyn = true;
        r.track = it->out_buf[0];
    }
    return r;
}

bool Planner::plan_graph(wae_graph* g, uint32_t gi) {
    Orderer ord{g};
    ord.run();
    if (!ord.broken.empty()) has_feedback = true;  // feedback through a DelayNode: its levels are replayed quantum by quantum
    // nodes from which a broken DelayWriter is reachable (reverse reachability over the ordered graph's edges, AudioParam ->
    // owner edges included): they have to be rendered quantum by quantum
    std::set<uint32_t> feeds_cycle;
    if (!ord.broken.empty()) {
        std::map<uint32_t, std::vector<uint32_t>> rev;
        for (uint32_t src = 0; src < (uint32_t)ord.edges.v.size(); src++)
            if (ord.edges.has[src])
                for (auto& e : *ord.edges.v[src]) rev[e.other_id].push_back(src);
        std::vector<uint32_t> todo(ord.broken.begin(), ord.broken.end());
        while (!todo.empty()) {
            uint32_t x = todo.back();
            todo.pop_back();
            if (!feeds_cycle.insert(x).second) continue;
            for (uint32_t y : rev[x]) todo.push_back(y);
        }
    }
    // ... and of those, the ones a cycle feeds (descendants of the readers of the broken delays): only they depend on
    // audio that is produced quantum by quantum
    std::set<uint32_t> fed_by_cycle;
    if (!ord.broken.empty()) {
        std::vector<uint32_t> todo;
        for (uint32_t w : ord.broken) todo.push_back(g->nodes.at(w).delay_peer);
        while (!todo.empty()) {
            uint32_t x = todo.back();
            todo.pop_back();
            if (!fed_by_cycle.insert(x).second) continue;
            const std::vector<Edge>* it = ord.edges.find(x);
            if (!it) continue;
            for (auto& e : *it) todo.push_back(e.other_id);
        }
    }
    auto node_class = [&](uint32_t id) {
        if (ord.broken.empty() || !feeds_cycle.count(id)) return ord.broken.empty() ? 0 : 2;
        return fed_by_cycle.count(id) ? 1 : 0;
    };
    NodeTable& pn = node_table;
    cur_pn = &pn;
    pn.reset(g->nodes.empty() ? 0 : g->nodes.max_id());
    for (auto& kv : g->nodes) pn.put(kv.first, &kv.second);
    // Graph::render (graph.rs:500-535): walk the order, append each audio edge to its destination port
    for (uint32_t id : ord.ordered) {
        Node& n = g->nodes.at(id);
        for (auto& e : ord.edges.at(id)) {
            if (e.other_index < 0) continue;
            PNode* it = pn.find(e.other_id);
            if (!it) continue;
            it->in_edges[e.other_index].push_back(PortRef{id, e.self_index});
        }
    }
    hm::SchedClock clock(g->sample_rate);
    const double sr = (double)g->sample_rate;
    // ---- chain fusion (WAE_OPT_FUSE): sources and biquad/gain/shaper nodes are not emitted one stage each; a node
    // with exactly one consumer stays PENDING, the consumer either extends the chain (same channel count, single
    // edge) or forces it to be materialised into an arena buffer.  A chain that ends at a destination whose only
    // input it is writes the final PCM directly.
    const bool fuse = eng->fuse;
    const bool want_scan_coefs = !dry || plan_digest_wanted();  // (the sizing pass needs their number only)
    std::map<uint32_t, PendingChain> pending;
    auto consumers = [&](const Node& nd) {
        int k = 0;
        for (auto& e : ord.edges.at(nd.id))
            if (e.other_index >= 0) k++;
        return k;
    };
    auto emit_chain = [&](PendingChain& pc, int L) {
        const int variant = pc.inst.src_kind * 6 + pc.inst.n_biquad * 2 + (pc.inst.has_shaper ? 1 : 0);
        const int consumer_cls = cur_cls;  // a chain is emitted while its consumer is planned, but runs with its own nodes' class
        cur_cls = pc.cls;
        StageBuild& cs = stage(L, S_CHAIN, variant);
        cur_cls = consumer_cls;
        for (int k = 0; k < pc.inst.n_biquad; k++) {
            pc.inst.bq[k].coef = cs.add_scan_coef(want_scan_coefs, [&] { return make_scan_coef(pc.coefs[k]); });
        }
        cs.max_ch = std::max(cs.max_ch, pc.ch);
        cs.chain.push_back(pc.inst);
    };
    // may_alias: the consumer reads its input through chan() with any alignment (the convolver's forward transform): a pending chain that
    // is nothing but an AudioBufferSourceNode playing its buffer 1:1 from frame 0, the buffer covering the whole (quantum-padded) render,
    // IS that buffer — no copy into the arena
    auto materialize = [&](uint32_t nid, bool may_alias = false) -> bool {
        auto it = pending.find(nid);
        if (it == pending.end()) return true;
        PNode& sp = pn.at(nid);
        {
            const ChainInst& ci = it->second.inst;
            const AbsnInst& a = ci.absn;
            bool unit = true;
            for (int i = 0; i < 4; i++) unit = unit && ci.g[i] == 1.f;
            if (may_alias && ci.src_kind == CHAIN_SRC_ABSN && ci.n_biquad == 0 && !ci.has_shaper && unit && it->second.phase == 0 &&
                !it->second.lay.dyn() && a.n_start == 0 && !a.loop && a.buf_offset == 0 && a.buf_len >= b->lq && a.buf_stride <= 0xffffffffll &&
                seg_start == 0 && seg_end >= b->lq) {
                sp.out_buf = {BufRef{const_cast<float*>(a.buf), (uint32_t)a.buf_stride, 1}};
                pending.erase(it);
                return true;
            }
        }
        BufRef buf = arena_buf(it->second.ch, it->second.lay.dyn());  // (k_chain writes the layout track itself)
        if (!buf.p) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
        it->second.inst.out = buf;
        it->second.inst.limit = -1;
        it->second.inst.out_dup = 0;
        emit_chain(it->second, 2 * sp.level + 1);
        sp.out_buf = {buf};
        pending.erase(it);
        return true;
    };
    // can a node of this kind still be appended to the canonical chain gain, A, gain, B, gain, shaper, gain?
    auto chain_accepts = [&](const PendingChain& pc, Kind kind) {
        if (kind == K_GAIN) return true;
        if (kind == K_BIQUAD) return pc.phase <= 1;
        if (kind == K_SHAPER) return pc.phase < 5;
        return false;
    };
    for (uint32_t id : ord.ordered) {
        Node& n = g->nodes.at(id);
        if (n.kind == K_LISTENER) continue;
        PNode& p = pn.at(id);
        cur_cls = node_class(id);
        key_graph = gi;
        key_node = id;
        key_seq = 0;
        key_salt = n.kind == K_PARAM ? (uint64_t)n.param.events.size() : 0;  // a param whose event list grew restarts its timeline
        if (n.kind == K_PARAM) {
            // AudioParamProcessor (param.rs:685-797): only params with automation events or audio-rate inputs become
            // GPU work; a constant param is a scalar in its owner's instance
            auto& edges = p.in_edges[0];
            // (a render without suspend points never replays a timeline: a constant param needs no record at all — most params are)
            if (seg_start == 0 && seg_end >= b->lq && edges.empty() && n.param.constant()) continue;
            const ParamTimeline* tlp = param_timeline(gi, id, n.param, g->sample_rate);
            if (n.param.constant() && edges.empty()) continue;
            int level = 0;
            for (auto& r : edges) level = std::max(level, pn.at(r.node).level + 1);
            p.level = level;
            for (auto& r : edges)
                if (!materialize(r.node)) return false;
            const ParamTimeline& tl = *tlp;
            if (!tl.error.empty()) return bail(WAE_NOT_SUPPORTED, tl.error);
            ParamInst pi{};
            if (!edges.empty()) {  // sum of the connected signals, first channel each (1 / explicit / discrete, param.rs:296-310)
                bool any_dyn = false;
                for (auto& r : edges) any_dyn = any_dyn || pn.at(r.node).lay_out(r.port).dyn();
                if (any_dyn) {  // edges whose layout changes: folded per quantum; a silent sum reads as zeros, which is what the param adds then
                    StageBuild& ms = stage(2 * level, S_MIX_DYN);
                    MixDynInst m{};
                    m.out = arena_buf(1);
                    if (!m.out.p) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                    m.out_ch = 1;
                    m.interp = WAE_INTERPRETATION_DISCRETE;
                    m.mode = WAE_COUNT_MODE_EXPLICIT;
                    m.cfg_count = 1;
                    m.n_edges = (int)edges.size();
                    m.edge_offset = (uint32_t)ms.mix_edges.size();
                    m.limit = -1;
                    for (auto& r : edges) ms.mix_edges.push_back(MixEdge{pn.at(r.node).out_buf[r.port], pn.at(r.node).out_ch[r.port], 0});
                    ms.mix_dyn.push_back(m);
                    pi.in = m.out;
                } else {
                    StageBuild& ms = stage(2 * level, S_MIX);
                    MixInst m{};
                    m.out = arena_buf(1);
                    if (!m.out.p) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                    m.out_ch = 1;
                    m.interp = WAE_INTERPRETATION_DISCRETE;
                    m.n_edges = (int)edges.size();
                    m.edge_offset = (uint32_t)ms.mix_edges.size();
                    m.limit = -1;
                    for (auto& r : edges) ms.mix_edges.push_back(MixEdge{pn.at(r.node).out_buf[r.port], pn.at(r.node).out_ch[r.port], 0});
                    ms.mix.push_back(m);
                    pi.in = m.out;
                }
            }
            pi.events = tl.events.empty() ? nullptr : upload(tl.events);
            pi.curves = tl.curves.empty() ? nullptr : upload(tl.curves);
            pi.state = alloc<ParamState>(1, true, true);
            pi.out = arena_buf(2);  // channel 0: value per frame, channel 1: single-valued flag per quantum
            if (!pi.state || !pi.out.p) return bail(WAE_OUT_OF_MEMORY, "out of device memory (param)");
            pi.def = n.param.default_value;
            pi.mn = n.param.min_value;
            pi.mx = n.param.max_value;
            pi.intrinsic0 = tl.intrinsic;
            pi.has_last0 = tl.has_last ? 1 : 0;
            pi.last0 = tl.last;
            pi.sample_rate = g->sample_rate;
            pi.n_events = (int32_t)tl.events.size();
            pi.a_rate = n.param.a_rate ? 1 : 0;
            stage(2 * level + 1, S_PARAM).param.push_back(pi);
            p.out_ch = {1};
            p.out_buf = {pi.out};
            continue;
        }
        // ---- inputs: static channel count + mix stage where needed
        int level = 0;
        for (auto& port : p.in_edges)
            for (auto& r : port) level = std::max(level, pn.at(r.node).level + 1);
        bool dyn_params = false;
        for (uint32_t pid : n.params) {
            PNode& pp = pn.at(pid);
            if (!pp.out_buf.empty()) {
                dyn_params = true;
                level = std::max(level, pp.level + 1);
            }
        }
        if (n.kind == K_PANNER)
            for (uint32_t pid = 2; pid <= 10; pid++)
                if (pn.count(pid) && !pn.at(pid).out_buf.empty()) level = std::max(level, pn.at(pid).level + 1);
        p.level = level;
        const bool fuse_n = fuse && !dyn_params;  // nodes with automated params run their own a-rate kernels
        // does this node extend the pending chain of its only producer / take it as the destination's only input?
        uint32_t fuse_src = 0;
        bool extend = false, dest_direct = false;
        const bool chain_kind = !dyn_params && ((n.kind == K_BIQUAD && !eng->serial_filters) || (fuse && (n.kind == K_GAIN || (n.kind == K_SHAPER && !(n.oversample && n.has_curve)))));
        if (fuse && n.n_inputs == 1 && p.in_edges[0].size() == 1 && p.in_edges[0][0].port == 0) {
            auto it = pending.find(p.in_edges[0][0].node);
            if (it != pending.end()) {
                int sch = pn.at(it->first).out_ch[0];
                if (chain_kind && computed_channels(n.cfg, sch) == sch && chain_accepts(it->second, n.kind)) {
                    extend = true;
                    fuse_src = it->first;
                } else if (n.kind == K_DEST && b->length <= 0xffffffffull &&
                           (sch == (int)b->channels || (sch == 1 && b->channels == 2 && n.cfg.interp == WAE_INTERPRETATION_SPEAKERS))) {
                    dest_direct = true;
                    fuse_src = it->first;
                }
            }
        }
        // ---- k_voice_sum (WAE_OPT_VOICE_SUM): a port fed by many oscillator -> [biquad] -> gain voices, all of them still pending chains
        // (mono, constant layout, one consumer): the voices are not materialised, one kernel renders them and keeps the running sum in
        // registers, in the port's edge order.  Only when the launch has enough (2048-frame tile, port) work items to fill the machine
        // about twice: one graph with thousands of voices and a short render (configs[2]) is better served by k_chain + k_mix, which
        // take their parallelism from the voices.
        std::vector<char> port_vsum(p.in_edges.size(), 0);
        std::vector<int> port_vsum_nb(p.in_edges.size(), 0);
        if (fuse && voice_sum_mode() != 0 && !extend && !dest_direct && cur_cls == 0 && n.kind != K_DELAY_R) {
            for (size_t pi = 0; pi < p.in_edges.size() && (int)pi < n.n_inputs; pi++) {
                const auto& edges = p.in_edges[pi];
                if ((int)edges.size() < 8) continue;
                const int ch = computed_channels(n.cfg, 1);
                if (!(ch == 1 || (ch == 2 && n.cfg.interp == WAE_INTERPRETATION_SPEAKERS))) continue;
                if (n.kind == K_DEST && b->length > 0xffffffffull) continue;
                const int64_t tiles = (seg_end - seg_start + 2047) / 2048;
                if (voice_sum_mode() < 2 && tiles * (int64_t)group_graphs < 2 * (int64_t)voice_sum_slots()) continue;
                int nb = -1;
                bool ok = true;
                std::set<uint32_t> seen_nodes;
                for (auto& r : edges) {
                    auto it = pending.find(r.node);
                    if (r.port != 0 || it == pending.end()) { ok = false; break; }
                    const PendingChain& pc = it->second;
                    if (pc.inst.src_kind != CHAIN_SRC_OSC || pc.ch != 1 || pc.inst.has_shaper || pc.inst.n_biquad > 1 || pc.lay.dyn() || pc.cls != cur_cls ||
                        (nb >= 0 && nb != pc.inst.n_biquad) || !seen_nodes.insert(r.node).second) { ok = false; break; }
                    nb = pc.inst.n_biquad;
                }
                if (!ok) continue;
                port_vsum[pi] = 1;
                port_vsum_nb[pi] = nb;
            }
        }
        for (size_t pi = 0; pi < p.in_edges.size(); pi++) {
            if (port_vsum[pi]) continue;
            auto& port = p.in_edges[pi];
            for (auto& r : port)
                if (!((extend || dest_direct) && r.node == fuse_src))
                    if (!materialize(r.node, n.kind == K_CONV && n.buffer && port.size() == 1)) return false;
        }
        p.in_ch.assign(n.n_inputs, 1);
        p.in_buf.assign(n.n_inputs, BufRef{nullptr, 0, 0});
        p.in_lay.assign(n.n_inputs, Lay::fixed(1));
        for (int port = 0; port < n.n_inputs; port++) {
            auto& edges = p.in_edges[port];
            int max_in = 1;
            for (auto& r : edges) max_in = std::max(max_in, pn.at(r.node).out_ch[r.port]);
            int ch = computed_channels(n.cfg, max_in);
            if (n.kind == K_DELAY_R) {  // the reader's only input is the hidden writer edge: take the writer's layout
                ch = max_in;
            }
            p.in_ch[port] = ch;
            p.in_lay[port] = Lay::fixed(ch);
            bool is_dest = n.kind == K_DEST;
            if (extend || dest_direct) {  // the producer's chain is consumed in registers / written directly
                if (extend) p.in_lay[port] = pending.at(fuse_src).lay;
                continue;
            }
            if (port_vsum[port]) {  // the voices of this port and their sum in one kernel
                StageBuild& vs = stage(2 * level, S_VSUM, port_vsum_nb[port]);
                VoiceGroup vg{};
                vg.first = (int32_t)vs.chain.size();
                vg.n_voices = (int32_t)edges.size();
                vg.out_dup = ch;
                vg.limit = -1;
                if (is_dest) {
                    vg.out = BufRef{b->d_out + (size_t)gi * b->channels * b->length, (uint32_t)b->length, 1};
                    vg.limit = (int64_t)b->length;
                } else {
                    vg.out = arena_buf(ch);
                    if (!vg.out.p) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                }
                for (auto& r : edges) {
                    PendingChain pc = std::move(pending.at(r.node));
                    pending.erase(r.node);
                    // (k_voice_sum prefetches the constants of voice k as coefficient set k: one set per voice, in voice order)
                    if (pc.inst.n_biquad == 1 && vs.n_scan_coef != vs.chain.size()) return bail(WAE_UNSUPPORTED, "internal: voice-sum coefficient table out of step");
                    for (int k = 0; k < pc.inst.n_biquad; k++)
                        pc.inst.bq[k].coef = vs.add_scan_coef(want_scan_coefs, [&] { return make_scan_coef(pc.coefs[k]); });
                    pc.inst.limit = -1;
                    pc.inst.out_dup = 0;
                    vs.chain.push_back(pc.inst);
                }
                vs.vgroups.push_back(vg);
                p.in_buf[port] = vg.out;
                continue;
            }
            if (is_dest && edges.size() == 1 && pn.at(edges[0].node).wrote_dest) {  // the producer already wrote the rendered PCM
                p.in_buf[port] = pn.at(edges[0].node).out_buf[edges[0].port];
                continue;
            }
            // ---- the port's layout over time: AudioRenderQuantum::add folded over the edges (quantum.rs:532-569)
            bool any_dyn = false;
            Lay pl = Lay::fixed(ch);
            if (!edges.empty()) {
                int lo = 1, hi = 1, on_nlo = 0, min_nlo = 255;
                bool all_may_silent = true;
                for (auto& r : edges) {
                    const Lay el = pn.at(r.node).lay_out(r.port);
                    any_dyn = any_dyn || el.dyn();
                    lo = std::max<int>(lo, el.lo);
                    hi = std::max<int>(hi, el.hi);
                    if (!el.may_silent) on_nlo = std::max<int>(on_nlo, el.nlo);
                    min_nlo = std::min<int>(min_nlo, el.nlo);
                    all_may_silent = all_may_silent && el.may_silent;
                }
                const int nlo = std::max(on_nlo, min_nlo);
                pl.lo = (uint8_t)computed_channels(n.cfg, lo);
                pl.hi = (uint8_t)computed_channels(n.cfg, hi);
                pl.nlo = (uint8_t)computed_channels(n.cfg, nlo);
                pl.nhi = pl.hi;
                pl.may_silent = all_may_silent;
                if (n.kind == K_DELAY_R) pl = pn.at(edges[0].node).lay_out(edges[0].port);
            }
            // more than two layouts meeting in a port wider than stereo: the order of the up-mixes matters (mono, stereo, 5.1: the
            // reference goes 1 -> 2 -> 6): fold edge by edge like it does
            bool needs_fold = false;
            if (ch > 2 && n.cfg.mode != WAE_COUNT_MODE_EXPLICIT)
                for (auto& r : edges) needs_fold = needs_fold || pn.at(r.node).out_ch[r.port] != ch;
            if (!is_dest && edges.size() == 1 && pn.at(edges[0].node).out_ch[edges[0].port] == ch) {
                const Lay el = pn.at(edges[0].node).lay_out(edges[0].port);
                // a single edge IS the port when computedNumberOfChannels leaves every count it can have alone
                const bool identity = !el.dyn() || n.kind == K_DELAY_R || n.cfg.mode == WAE_COUNT_MODE_MAX ||
                                      (n.cfg.mode == WAE_COUNT_MODE_CLAMPED_MAX && el.hi <= n.cfg.count);
                // the time-batched convolver reads all static channels of every quantum: it needs the canonical PCM k_mix_dyn writes
                const bool canonical_needed = el.dyn() && n.kind == K_CONV;
                if (identity && !canonical_needed) {
                    p.in_buf[port] = pn.at(edges[0].node).out_buf[edges[0].port];  // alias, no copy
                    p.in_lay[port] = el;
                    continue;
                }
            }
            if (any_dyn || needs_fold) {
                StageBuild& ms = stage(2 * level, S_MIX_DYN);
                MixDynInst m{};
                m.out_ch = ch;
                m.interp = n.cfg.interp;
                m.mode = n.cfg.mode;
                m.cfg_count = n.cfg.count;
                m.n_edges = (int)edges.size();
                m.edge_offset = (uint32_t)ms.mix_edges.size();
                m.limit = -1;
                if (is_dest) {
                    m.out = BufRef{b->d_out + (size_t)gi * b->channels * b->length, (uint32_t)b->length, 1};
                    m.limit = (int64_t)b->length;
                    if (b->length > 0xffffffffull) return bail(WAE_UNSUPPORTED, "render length above 2^32 frames");
                } else {
                    m.out = arena_buf(ch, pl.dyn());
                    if (!m.out.p) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                }
                for (auto& r : edges) {
                    PNode& sn = pn.at(r.node);
                    ms.mix_edges.push_back(MixEdge{sn.out_buf[r.port], sn.out_ch[r.port], 0});
                }
                ms.mix_dyn.push_back(m);
                p.in_buf[port] = m.out;
                p.in_lay[port] = pl;
                continue;
            }
            StageBuild& ms = stage(2 * level, S_MIX);
            MixInst m{};
            m.out_ch = ch;
            m.interp = n.cfg.interp;
            m.n_edges = (int)edges.size();
            m.edge_offset = (uint32_t)ms.mix_edges.size();
            m.limit = -1;
            if (is_dest) {
                m.out = BufRef{b->d_out + (size_t)gi * b->channels * b->length, (uint32_t)b->length, 1};
                m.limit = (int64_t)b->length;
                if (b->length > 0xffffffffull) return bail(WAE_UNSUPPORTED, "render length above 2^32 frames");
            } else {
                m.out = arena_buf(ch);
                if (!m.out.p) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
            }
            for (auto& r : edges) {
                PNode& s = pn.at(r.node);
                ms.mix_edges.push_back(MixEdge{s.out_buf[r.port], s.out_ch[r.port], 0});
            }
            ms.mix.push_back(m);
            p.in_buf[port] = m.out;
        }
        const int L = 2 * level + 1;  // node kernels run after the mixes of their level
        auto need_out = [&](int ch) {
            p.out_ch = {ch};
            p.out_buf = {arena_buf(ch)};
            return p.out_buf[0].p != nullptr;
        };
        // the node's (single) output has a layout that is not constant: give its buffer a layout track
        auto out_dynamic = [&](const Lay& l) {
            p.out_lay = {l};
            if (l.dyn() && !p.out_buf.empty() && p.out_buf[0].p && !p.out_buf[0].absolute) {
                p.out_buf[0].meta = dry ? reinterpret_cast<uint8_t*>(uintptr_t(256)) : reinterpret_cast<uint8_t*>(p.out_buf[0].p + (size_t)p.out_ch[0] * (size_t)b->chunk);
                p.out_buf[0].meta_stride = (uint32_t)((b->chunk / 128 + 16) / 16 * 16);
            }
        };
        // a scheduled source: `ch` channels inside [n_first, n_stop), one silent channel outside (never silent when it covers the render)
        auto source_lay = [&](int64_t n_first, int64_t n_stop, int ch) { return (n_first <= 0 && n_stop >= b->lq) ? Lay::fixed(ch) : Lay::gated(ch); };
        auto source_meta = [&](int64_t n_first, int64_t n_stop, int ch) {
            MetaInst m{};
            m.out = p.out_buf[0];
            m.mode = META_SOURCE;
            m.out_ch = ch;
            m.count = ch;
            m.n_first = n_first;
            m.n_stop = n_stop;
            stage(L, S_META).meta.push_back(m);
        };
        const Lay in0 = p.in_lay.empty() ? Lay::fixed(1) : p.in_lay[0];
        // biquad / IIR (biquad_filter.rs:778-815): silent once the input is and the tail has rung out; keeps the channels of the last
        // input that was not silent
        auto filter_lay = [](const Lay& l) { return Lay{(uint8_t)(l.may_silent ? 1 : l.lo), l.hi, l.nlo, l.nhi, l.may_silent}; };
        // output with the input's layout and channel count: share the input's layout track
        auto out_like_input = [&]() {
            p.out_lay = {in0};
            if (in0.dyn() && !p.out_buf.empty() && p.out_buf[0].p) {
                p.out_buf[0].meta = p.in_buf[0].meta;
                p.out_buf[0].meta_stride = p.in_buf[0].meta_stride;
            }
        };
        // registers this node as the tail of a chain: pending while exactly one consumer may still fuse with it
        auto finish_chain = [&](PendingChain&& pc) -> bool {
            p.out_ch = {pc.ch};
            p.out_buf = {BufRef{nullptr, 0, 0}};
            p.out_lay = {pc.lay};
            pending[id] = std::move(pc);
            if (!(fuse && consumers(n) == 1)) return materialize(id);
            return true;
        };
        auto source_chain = [&](int kind, int ch) {
            PendingChain pc;
            std::memset(&pc.inst, 0, sizeof(pc.inst));
            pc.inst.src_kind = kind;
            pc.inst.ch = ch;
            pc.inst.limit = -1;
            for (int i = 0; i < 4; i++) pc.inst.g[i] = 1.f;
            pc.ch = ch;
            pc.cls = cur_cls;
            pc.lay = Lay::fixed(ch);
            return pc;
        };
        // chain that this biquad / gain / shaper node joins: its producer's pending chain, or a new one reading in_buf
        auto open_chain = [&]() {
            if (extend) {
                PendingChain pc = std::move(pending.at(fuse_src));
                pending.erase(fuse_src);
                return pc;
            }
            PendingChain pc = source_chain(CHAIN_SRC_BUFFER, p.in_ch[0]);
            pc.inst.in = p.in_buf[0];
            pc.lay = in0;
            return pc;
        };
        switch (n.kind) {
            case K_DEST: {
                p.out_ch = {(int)b->channels};
                if (dest_direct) {  // the chain writes the rendered PCM itself (speaker up-mix 1->2 = copy, quantum.rs:301-305)
                    PendingChain pc = std::move(pending.at(fuse_src));
                    pending.erase(fuse_src);
                    BufRef fin{b->d_out + (size_t)gi * b->channels * b->length, (uint32_t)b->length, 1};
                    pc.inst.out = fin;
                    pc.inst.limit = (int64_t)b->length;
                    pc.inst.out_dup = (pc.ch == 1 && b->channels == 2) ? 2 : 0;
                    emit_chain(pc, L);
                    pn.at(fuse_src).out_buf = {fin};
                    p.in_buf[0] = fin;
                }
                p.out_buf = {p.in_buf[0]};
                algorithmic_bytes += (uint64_t)b->channels * b->length * 4;  // destination write, SURVEY §8(d)
                break;
            }
            case K_OSC: {
                PRef pf = param_ref(g, n.params[0]), pd = param_ref(g, n.params[1]);
                float freq = pf.v, detune = pd.v;
                if (!fuse_n && !need_out(1)) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                OscInst o{};
                double start_ratio = 0.;
                if (!fuse_n) o.out = p.out_buf[0];
                o.type = n.type;
                double computed_freq = (double)freq * std::exp2((double)detune / 1200.);  // oscillator.rs:30-32
                o.incr = computed_freq / sr;
                o.inv_incr = o.incr != 0. ? 1. / o.incr : 0.;
                o.outside_nyquist = std::fabs(computed_freq) >= sr / 2.;
                o.n_first = std::numeric_limits<int64_t>::max();
                o.n_stop = std::numeric_limits<int64_t>::max();
                o.phase0 = 0.;
                if (n.start_time < 1e300) {
                    // oscillator.rs:391-428,511-540: first rendered frame and its phase
                    int64_t q = clock.quantum_containing(n.start_time);
                    double start = n.start_time;
                    if (start < clock.block_time(q)) start = clock.block_time(q);  // "prevent scheduling in the past"
                    double t = 0.;
                    // walk the accumulated per-frame clock of that quantum
                    double cur = clock.block_time(q);
                    int i = 0;
                    for (; i < 128; i++) {
                        if (!(cur < start)) break;
                        cur += clock.dt;
                    }
                    t = cur;
                    o.n_first = q * 128 + i;
                    if (i < 128 && t > start) {
                        double ratio = (t - start) / clock.dt;
                        start_ratio = ratio;
                        double ph = o.incr * ratio;
                        if (o.outside_nyquist) {
                            ph = std::fmod(ph, 1.);
                            if (ph < 0.) ph += 1.;
                        } else {
                            ph = hm::unroll_phase(ph);
                        }
                        o.phase0 = ph;
                    }
                    if (n.stop_time < 1e300) {
                        int64_t qs = clock.quantum_containing(n.stop_time);
                        if (n.stop_time <= clock.block_time(qs)) o.n_stop = qs * 128;
                        else o.n_stop = clock.first_frame_at_or_after(n.stop_time);
                    }
                }
                if (n.type == WAE_OSC_CUSTOM) {
                    float* d = upload(n.table);
                    o.table = d;
                    o.table_len = (int)n.table.size();
                } else {
                    o.table = eng->d_sine;
                    o.table_len = 2048;
                }
                o.fast = (!o.outside_nyquist && o.incr > 0. && o.incr < 0.5 && (o.table_len == 2048 || (o.type != WAE_OSC_SINE && o.type != WAE_OSC_CUSTOM))) ? 1 : 0;
                if (dyn_params) {  // automated / audio-rate frequency or detune: running-sum phase
                    OscArInst oa{};
                    oa.base = o;
                    oa.freq = pf.dyn ? pf.track : BufRef{nullptr, 0, 0};
                    oa.detune = pd.dyn ? pd.track : BufRef{nullptr, 0, 0};
                    oa.f_val = freq;
                    oa.d_val = detune;
                    oa.start_ratio = start_ratio;
                    oa.phase = alloc<double>(1, true, true);
                    oa.sample_rate = g->sample_rate;
                    if (!oa.phase) return bail(WAE_OUT_OF_MEMORY, "out of device memory (state)");
                    out_dynamic(source_lay(o.n_first, o.n_stop, 1));
                    oa.base.out = p.out_buf[0];
                    if (p.out_lay[0].dyn()) source_meta(o.n_first, o.n_stop, 1);
                    stage(L, S_OSC_AR).osc_ar.push_back(oa);
                } else if (fuse_n) {
                    PendingChain pc = source_chain(CHAIN_SRC_OSC, 1);
                    pc.inst.osc = o;
                    pc.lay = source_lay(o.n_first, o.n_stop, 1);
                    if (!finish_chain(std::move(pc))) return false;
                } else {
                    out_dynamic(source_lay(o.n_first, o.n_stop, 1));
                    o.out = p.out_buf[0];
                    if (p.out_lay[0].dyn()) source_meta(o.n_first, o.n_stop, 1);
                    stage(L, S_OSC).osc.push_back(o);
                }
                break;
            }
            case K_CONST: {
                PRef po = param_ref(g, n.params[0]);
                float v = po.v;
                if (!fuse_n && !need_out(1)) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                ConstInst c{};
                if (!fuse_n) c.out = p.out_buf[0];
                if (po.dyn) c.track = po.track;
                c.value = v;
                c.n_first = std::numeric_limits<int64_t>::max();
                c.n_stop = std::numeric_limits<int64_t>::max();
                if (n.start_time < 1e300) {
                    // constant_source.rs:203-246
                    c.n_first = clock.first_frame_at_or_after(n.start_time);
                    if (n.stop_time < 1e300) c.n_stop = clock.first_frame_at_or_after(n.stop_time);
                }
                if (fuse_n) {
                    PendingChain pc = source_chain(CHAIN_SRC_CONST, 1);
                    pc.inst.cst = c;
                    pc.lay = source_lay(c.n_first, c.n_stop, 1);
                    if (!finish_chain(std::move(pc))) return false;
                } else {
                    out_dynamic(source_lay(c.n_first, c.n_stop, 1));
                    c.out = p.out_buf[0];
                    if (p.out_lay[0].dyn()) source_meta(c.n_first, c.n_stop, 1);
                    stage(L, S_CONST).cst.push_back(c);
                }
                break;
            }
            case K_ABSN: {
                PRef pdet = param_ref(g, n.params[0]), prate = param_ref(g, n.params[1]);
                const float detune = pdet.v, rate = prate.v;
                const bool rate_automated = pdet.dyn || prate.dyn;
                int ch = n.buffer ? (int)n.buffer->channels.size() : 1;
                if (!n.buffer || n.start_time >= 1e300 || ch == 0) {  // never plays: silence
                    if (!need_out(1)) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                    StageBuild& ms = stage(L, S_MIX);
                    ms.mix.push_back(MixInst{p.out_buf[0], 1, 0, 0, (uint32_t)ms.mix_edges.size(), -1});
                    out_dynamic(Lay{1, 1, 1, 1, true});  // silent for good
                    if (p.out_buf[0].meta) source_meta(0, 0, 1);
                    break;
                }
                PcmBuffer& pb = *n.buffer;
                double computed_rate = (double)rate * std::exp2((double)detune / 1200.);
                double duration = pb.duration();
                double ls = n.loop_start, le = n.loop_end;  // clamp_loop_boundaries, audio_buffer_source.rs:400-417
                if (ls < 0.) ls = 0.; else if (ls > duration) ls = duration;
                if (le <= 0. || le > duration) le = duration;
                int64_t q = clock.quantum_containing(n.start_time);
                // a start time that IS the next block boundary but compares below next_block_time by one rounding:
                // the reference goes through one all-silent slow-track quantum, then aligns (audio_buffer_source.rs:521-523)
                if (n.start_time > clock.block_time(q) && n.start_time == clock.block_time(q + 1)) q = q + 1;
                bool aligned = (n.start_time <= clock.block_time(q)) && n.offset == 0.;  // start in the past snaps to the block
                bool fast = !rate_automated && aligned && (double)pb.sample_rate / sr == 1. && computed_rate == 1. && ls == 0. && le == duration &&
                            n.duration > 1e300 && n.stop_time > 1e300;
                // everything the closed-form tracks do not cover runs the renderer's own frame loop (one warp per source)
                bool serial = rate_automated || (!fast && !(computed_rate > 0.));
                if (!fast && !serial && n.loop) {
                    const bool custom = ls >= 0. && le > 0. && ls < le;
                    const double loop_len = custom ? le - ls : duration;
                    if (!(loop_len > 4. * clock.dt * computed_rate)) serial = true;  // loop shorter than four output frames
                }
                const bool fuse_src = fuse_n && fast;
                if (!fuse_src && !need_out(ch)) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                size_t len = pb.length();
                size_t stride = (len + 3) / 4 * 4;  // every channel starts 16 B aligned (LDG.128)
                // one copy of the PCM per AudioBuffer in the group's slab (the grains of a granular patch all play the same one: the graph
                // holds it once, wae_abi_graph.cpp copy_buffer), shared by the plans of all render segments
                auto so = src_offsets.find({gi, id});
                bool first_use = so == src_offsets.end();
                if (first_use) {
                    auto bo = buf_offsets.find(n.buffer.get());
                    if (bo != buf_offsets.end()) first_use = false;  // (already in the slab for another node)
                    else bo = buf_offsets.emplace(n.buffer.get(), src_cursor).first;
                    so = src_offsets.emplace(std::make_pair(gi, id), bo->second).first;
                }
                float* d_buf = d_src + so->second;
                if (first_use) {
                    if (src_copies)  // recorded by the sizing pass: uploaded straight from the graph's buffer, planar [ch][stride] like the slab
                        src_copies->push_back(wae_batch::Group::SrcCopy{n.buffer, src_cursor, (size_t)ch * stride});
                    src_cursor += (size_t)ch * stride;
                    b->asset_bytes += (size_t)ch * len * 4;
                }
                if (serial) {
                    AbsnSerialInst a{};
                    a.out = p.out_buf[0];
                    a.buf = d_buf;
                    a.buf_len = (int64_t)len;
                    a.buf_stride = (int64_t)stride;
                    a.start_time = n.start_time;
                    a.stop_time = n.stop_time;
                    a.offset = n.offset;
                    a.duration = n.duration;
                    a.loop_start = ls;
                    a.loop_end = le;
                    a.buffer_duration = duration;
                    a.buffer_sample_rate = (double)pb.sample_rate;
                    a.sample_rate = sr;
                    const BufRef none{nullptr, 0, 0};
                    a.rate_track = prate.dyn ? prate.track : none;
                    a.detune_track = pdet.dyn ? pdet.track : none;
                    a.rate = rate;
                    a.detune = detune;
                    a.ch = ch;
                    a.loop = n.loop ? 1 : 0;
                    a.state = alloc<AbsnSerialState>(1, true, true);
                    if (!a.state) return bail(WAE_OUT_OF_MEMORY, "out of device memory (buffer source state)");
                    out_dynamic(Lay::gated(ch));  // when it plays depends on the automated rate: the kernel writes the layout track
                    a.out = p.out_buf[0];
                    stage(L, S_ABSN_SERIAL).absn_serial.push_back(a);
                    algorithmic_bytes += (uint64_t)ch * 4ull * (uint64_t)std::min<int64_t>(b->lq, (int64_t)len);
                    break;
                }
                if (!fast) {
                    // ---- slow track (audio_buffer_source.rs:625-823): fractional playhead
                    auto almost_equal = [](double x, double y) {
                        if (x == y) return true;
                        const double tol = 1.4901161193847656e-8;
                        double d = std::fabs(y - x);
                        return d <= tol || d <= std::max(std::fabs(x), std::fabs(y)) * tol;
                    };
                    AbsnSlowInst a{};
                    a.out = p.out_buf[0];
                    a.buf = d_buf;
                    a.buf_len = (int64_t)len;
                    a.buf_stride = (int64_t)stride;
                    a.ch = ch;
                    a.loop = n.loop ? 1 : 0;
                    a.sample_rate = sr;
                    a.buffer_duration = duration;
                    a.pos_scale = ((double)pb.sample_rate / sr) * sr;  // position = buffer_time * sampling_ratio; playhead = position * sr
                    a.step = clock.dt * computed_rate;
                    a.duration = n.duration;
                    // actual loop points (:627-636)
                    if (n.loop && ls >= 0. && le > 0. && ls < le) {
                        a.loop_start = ls;
                        a.loop_end = le;
                    } else {
                        a.loop_start = 0.;
                        a.loop_end = duration;
                    }
                    // first frame at / after the start time: current_time = block_time + i * dt (:648), sticky within
                    // almost::equal (:652-654)
                    double start = n.start_time;
                    int64_t qq = clock.quantum_containing(start);
                    int64_t n_first = -1;
                    double t_first = 0.;
                    for (int guard = 0; guard < 3 && n_first < 0; guard++, qq++) {
                        double bt0 = clock.block_time(qq);
                        for (int i = 0; i < 128; i++) {
                            double t = bt0 + (double)i * clock.dt;
                            if (almost_equal(t, start)) start = t;
                            if (!(t < start)) {
                                n_first = qq * 128 + i;
                                t_first = t;
                                break;
                            }
                        }
                    }
                    if (n_first < 0) n_first = qq * 128;
                    double delta = t_first - start;
                    double off = n.offset + delta * computed_rate;  // :672-674
                    off = std::min(std::max(off, 0.), duration);
                    if (n.loop && off > a.loop_end) off = a.loop_end;  // :676-678 (rate >= 0)
                    a.offset0 = off;
                    a.elapsed0 = std::fabs(delta * computed_rate);
                    a.n_first = n_first;
                    a.n_stop = std::numeric_limits<int64_t>::max();
                    if (n.stop_time < 1e300) {  // first frame with current_time >= stop_time (:663)
                        int64_t qs = clock.quantum_containing(n.stop_time);
                        int64_t ns = (qs + 1) * 128;
                        double bt0 = clock.block_time(qs);
                        for (int i = 0; i < 128; i++)
                            if (bt0 + (double)i * clock.dt >= n.stop_time) {
                                ns = qs * 128 + i;
                                break;
                            }
                        a.n_stop = ns;
                    }
                    // playhead schedule: walk the reference's per-frame bookkeeping (:730-770) from event to event — a
                    // frame where buffer_time is snapped to a loop point (almost::equal) or wrapped starts a new segment
                    std::vector<int64_t> seg_n{n_first};
                    std::vector<double> seg_bt{off};
                    if (n.loop && off < a.loop_end) {
                        const double ls2 = a.loop_start, le2 = a.loop_end, len2 = le2 - ls2, step = a.step;
                        const int64_t n_end = std::min<int64_t>(b->lq, a.n_stop);
                        int64_t m = 0;   // frames since n_first
                        double v = off;  // buffer_time of frame m
                        bool entered = false;
                        auto tz = [&](double x) { return 3.0e-8 * (1.0 + std::fabs(x)); };  // a little wider than almost::equal
                        while (n_first + m < n_end) {
                            // frames until the playhead can touch the tolerance zone of a loop point
                            double to_ls = v < ls2 - tz(ls2) ? (ls2 - tz(ls2) - v) / step : 0.;
                            double to_le = v < le2 - tz(le2) ? (le2 - tz(le2) - v) / step : 0.;
                            double skip = (!entered && to_ls > 0.) ? std::min(to_ls, to_le) : to_le;
                            int64_t adv = (int64_t)std::floor(skip);
                            if (adv > 0) {
                                v += (double)adv * step;
                                m += adv;
                                continue;
                            }
                            // exact per-frame logic of the reference
                            double w = v;
                            if (almost_equal(w, le2)) w = le2;
                            if (almost_equal(w, ls2)) w = ls2;
                            if (!entered && w >= ls2) entered = true;
                            if (entered) {
                                while (w >= le2) w -= len2;
                                while (w < ls2) w += len2;
                            }
                            if (w != v && n_first + m > seg_n.back()) {
                                seg_n.push_back(n_first + m);
                                seg_bt.push_back(w);
                            } else if (w != v) {
                                seg_bt.back() = w;
                            }
                            v = w + step;
                            m += 1;
                        }
                    }
                    a.n_seg = (int32_t)seg_n.size();
                    a.seg_n = upload(seg_n);
                    a.seg_bt = upload(seg_bt);
                    {
                        // layout: silent before the quantum of the first playing frame and after the quantum in which the source ends
                        // (stop time, explicit duration, or — not looping — the end of the buffer; audio_buffer_source.rs:826-838)
                        int64_t n_end = a.n_stop;
                        if (a.step > 0.) {
                            if (!n.loop) n_end = std::min<int64_t>(n_end, n_first + (int64_t)std::ceil(std::max(0., duration - off) / a.step));
                            if (n.duration < 1e300) n_end = std::min<int64_t>(n_end, n_first + (int64_t)std::ceil(std::max(0., n.duration - a.elapsed0) / a.step));
                        }
                        out_dynamic(source_lay(n_first, n_end, ch));
                        a.out = p.out_buf[0];
                        if (p.out_lay[0].dyn()) source_meta(n_first, n_end, ch);
                    }
                    stage(L, S_ABSN_SLOW).absn_slow.push_back(a);
                    algorithmic_bytes += (uint64_t)ch * 4ull * (uint64_t)std::min<int64_t>(b->lq, (int64_t)len);
                    break;
                }
                AbsnInst a{};
                if (!fuse_src) a.out = p.out_buf[0];
                a.buf = d_buf;
                a.buf_len = (int64_t)len;
                a.buf_stride = (int64_t)stride;
                a.n_start = q * 128;
                a.n_stop = std::numeric_limits<int64_t>::max();
                a.buf_offset = 0;
                a.ch = ch;
                a.loop = n.loop ? 1 : 0;
                if (!n.loop) {
                    // the quantum after which the source has `ended`: the reference accumulates buffer_time += block_duration and stops
                    // once it reaches the buffer's duration (audio_buffer_source.rs:609,826-838) — replayed, not divided
                    const double block_duration = clock.dt * 128.;
                    const int64_t max_q = (b->lq - a.n_start) / 128 + 2;
                    int64_t played = 0;
                    double bt = 0.;
                    while (played < max_q) {
                        bt += block_duration;
                        played++;
                        if (bt >= duration) break;
                    }
                    a.n_stop = a.n_start + played * 128;
                }
                if (fuse_src) {
                    PendingChain pc = source_chain(CHAIN_SRC_ABSN, ch);
                    pc.inst.absn = a;
                    pc.lay = source_lay(a.n_start, a.n_stop, ch);
                    if (!finish_chain(std::move(pc))) return false;
                } else {
                    out_dynamic(source_lay(a.n_start, a.n_stop, ch));
                    a.out = p.out_buf[0];
                    if (p.out_lay[0].dyn()) source_meta(a.n_start, a.n_stop, ch);
                    stage(L, S_ABSN).absn.push_back(a);
                }
                // compulsory read of the source PCM that is actually played
                algorithmic_bytes += (uint64_t)ch * 4ull * (uint64_t)std::max<int64_t>(0, std::min<int64_t>(b->lq - a.n_start, n.loop ? b->lq : (int64_t)len));
                break;
            }
            case K_BIQUAD: {
                PRef pq = param_ref(g, n.params[0]), pdt = param_ref(g, n.params[1]), pfr = param_ref(g, n.params[2]), pg = param_ref(g, n.params[3]);
                float q = pq.v, detune = pdt.v, freq = pfr.v, gain = pg.v;
                int ch = p.in_ch[0];
                if (dyn_params) {  // per-frame coefficients (biquad_filter.rs:837-855): serial a-rate kernel
                    if (!need_out(ch)) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                    BiquadArInst ba{};
                    ba.in = p.in_buf[0];
                    ba.out = p.out_buf[0];
                    const BufRef none{nullptr, 0, 0};
                    ba.q = pq.dyn ? pq.track : none;
                    ba.detune = pdt.dyn ? pdt.track : none;
                    ba.freq = pfr.dyn ? pfr.track : none;
                    ba.gain = pg.dyn ? pg.track : none;
                    ba.q_val = q; ba.detune_val = detune; ba.freq_val = freq; ba.gain_val = gain;
                    ba.state = alloc<double>((size_t)ch * 4, true, true);
                    if (!ba.state) return bail(WAE_OUT_OF_MEMORY, "out of device memory (state)");
                    if (in0.dyn()) {
                        ba.dyn_len = alloc<int32_t>((size_t)ch, true, true);
                        if (!ba.dyn_len) return bail(WAE_OUT_OF_MEMORY, "out of device memory (state)");
                        out_dynamic(filter_lay(in0));
                        ba.out = p.out_buf[0];
                    }
                    {
                        // five planes of f64 coefficients per frame of the chunk (an arena buffer of 10 float channels read as doubles)
                        BufRef cb = arena_buf(10);
                        if (!cb.p) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                        cb.stride = (uint32_t)b->chunk;  // in DOUBLES: plane k starts at double index k * chunk
                        ba.coefs = cb;
                    }
                    ba.sample_rate = g->sample_rate;
                    ba.type = n.type;
                    ba.ch = ch;
                    StageBuild& sb = stage(L, S_BIQUAD_AR);
                    sb.max_ch = std::max(sb.max_ch, ch);
                    sb.biquad_ar.push_back(ba);
                    break;
                }
                float cf = hm::biquad_computed_freq(freq, detune);
                hm::BiquadCoefs c = hm::biquad_coefs(n.type, sr, (double)cf, (double)gain, (double)q);
                double* state = alloc<double>((size_t)ch * 4, true, true);
                if (!state) return bail(WAE_OUT_OF_MEMORY, "out of device memory (state)");
                // An input whose channel COUNT changes while it sounds resets / drops channels of the filter mid-render
                // (biquad_filter.rs:798-815): the serial kernel follows it quantum by quantum; the scan keeps one state per channel
                const bool count_varies = !extend && in0.dyn() && !(in0.nlo == in0.nhi && in0.nhi == ch);
                if (eng->serial_filters || count_varies) {  // bit-faithful serial recurrence, one stage per biquad
                    if (!need_out(ch)) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                    BiquadInst bi{};
                    if (in0.dyn()) {
                        bi.dyn_len = alloc<int32_t>((size_t)ch, true, true);
                        if (!bi.dyn_len) return bail(WAE_OUT_OF_MEMORY, "out of device memory (state)");
                        out_dynamic(filter_lay(in0));
                    }
                    bi.in = p.in_buf[0];
                    bi.out = p.out_buf[0];
                    bi.b0 = c.b0; bi.b1 = c.b1; bi.b2 = c.b2; bi.a1 = c.a1; bi.a2 = c.a2;
                    bi.ch = ch;
                    bi.state = state;
                    StageBuild& s = stage(L, S_BIQUAD);
                    s.max_ch = std::max(s.max_ch, ch);
                    s.biquad.push_back(bi);
                    break;
                }
                PendingChain pc = open_chain();
                ChainBiquad& st = pc.inst.bq[pc.inst.n_biquad++];
                st.state = state;
                st.b0 = c.b0; st.b1 = c.b1; st.b2 = c.b2; st.a1 = c.a1; st.a2 = c.a2;
                pc.coefs[pc.inst.n_biquad - 1] = c;
                pc.phase = pc.phase == 0 ? 1 : 3;
                pc.lay = filter_lay(pc.lay);
                if (!finish_chain(std::move(pc))) return false;
                break;
            }
            case K_IIR: {
                int ch = p.in_ch[0];
                std::vector<double> ff = n.feedforward, fb = n.feedback;  // iir_filter.rs:282-309
                if (ff.size() < fb.size()) ff.resize(fb.size(), 0.);
                if (ff.size() > fb.size()) fb.resize(ff.size(), 0.);
                if (ff.size() <= 3 && !eng->serial_filters && !in0.dyn() && seg_start == 0 && seg_end >= b->lq) {
                    // Order <= 2 with a constant input layout: the same transfer function as a biquad — rendered by the time-parallel scan
                    // of k_chain (direct form I there, transposed direct form II in iir_filter.rs:386-407: the outputs differ in the last
                    // bits of the f64 arithmetic only) instead of one serial thread per channel.  With an input that can fall silent the
                    // serial kernel stays: its tail test looks at the reference's own state variables.  Same for a render cut by suspend
                    // points: the two forms keep different state (x / y history here, the reference's 20 accumulators there) and a later
                    // segment may see a layout that needs the serial kernel — the filter memory must survive the cut (fuzz seeds 31, 33).
                    ff.resize(3, 0.);
                    fb.resize(3, 0.);
                    const double a0 = fb[0];
                    hm::BiquadCoefs c{ff[0] / a0, ff[1] / a0, ff[2] / a0, fb[1] / a0, fb[2] / a0};
                    double* state = alloc<double>((size_t)ch * 4, true, true);
                    if (!state) return bail(WAE_OUT_OF_MEMORY, "out of device memory (state)");
                    PendingChain pc = open_chain();
                    ChainBiquad& st = pc.inst.bq[pc.inst.n_biquad++];
                    st.state = state;
                    st.b0 = c.b0; st.b1 = c.b1; st.b2 = c.b2; st.a1 = c.a1; st.a2 = c.a2;
                    pc.coefs[pc.inst.n_biquad - 1] = c;
                    pc.phase = 1;
                    if (!finish_chain(std::move(pc))) return false;
                    break;
                }
                if (!need_out(ch)) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                IirInst ii{};
                ii.in = p.in_buf[0];
                ii.out = p.out_buf[0];
                ii.n = (int)ff.size();
                ii.ch = ch;
                double a0 = fb[0];
                for (size_t i = 0; i < ff.size(); i++) {
                    ii.b[i] = ff[i] / a0;
                    ii.a[i] = fb[i] / a0;
                }
                ii.state = alloc<double>((size_t)ch * 20, true, true);
                if (!ii.state) return bail(WAE_OUT_OF_MEMORY, "out of device memory (state)");
                if (in0.dyn()) {
                    ii.dyn_len = alloc<int32_t>((size_t)ch, true, true);
                    if (!ii.dyn_len) return bail(WAE_OUT_OF_MEMORY, "out of device memory (state)");
                    out_dynamic(filter_lay(in0));
                    ii.out = p.out_buf[0];
                }
                StageBuild& s = stage(L, S_IIR);
                s.max_ch = std::max(s.max_ch, ch);
                s.iir.push_back(ii);
                break;
            }
            case K_GAIN: {
                PRef pgn = param_ref(g, n.params[0]);
                float gv = pgn.v;
                int ch = p.in_ch[0];
                if (pgn.dyn) {  // a-rate gain (gain.rs:189-197)
                    if (!need_out(ch)) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                    out_like_input();  // silent in -> silent out (gain.rs:155-158); the ~0 / ~1 shortcuts only exist for single values
                    stage(L, S_GAIN).gain.push_back(GainInst{p.in_buf[0], p.out_buf[0], gv, ch, pgn.track});
                    break;
                }
                // gain.rs:153-169: |g| <= 1e-6 -> silence, |1-g| <= 1e-6 -> pass-through (quanta >= 1; quantum 0 takes
                // the multiply path, a difference of at most 1e-6 * |x| that is below the parity tolerance)
                if (std::fabs(gv) <= 1e-6f) gv = 0.f;
                else if (std::fabs(1.f - gv) <= 1e-6f) gv = 1.f;
                if (fuse_n) {
                    PendingChain pc = open_chain();
                    // consecutive gains of one slot are folded (differs from two f32 multiplies by <= 1 ulp)
                    pc.inst.g[pc.phase == 0 ? 0 : (pc.phase == 1 ? 1 : (pc.phase == 3 ? 2 : 3))] *= gv;
                    if (gv == 0.f) pc.lay = Lay{1, 1, 1, 1, true};  // a gain of (about) zero answers with silence (gain.rs:160-163)
                    if (!finish_chain(std::move(pc))) return false;
                    break;
                }
                if (!need_out(ch)) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                if (gv == 0.f) {
                    out_dynamic(Lay{1, 1, 1, 1, true});
                    if (p.out_buf[0].meta) source_meta(0, 0, 1);
                } else {
                    out_like_input();
                }
                stage(L, S_GAIN).gain.push_back(GainInst{p.in_buf[0], p.out_buf[0], gv, ch, BufRef{nullptr, 0, 0}});
                break;
            }
            case K_SHAPER: {
                int ch = p.in_ch[0];
                const float* curve = n.has_curve ? upload(n.table) : nullptr;
                // can_propagate_silence (waveshaper.rs:480-503): the curve maps 0 to 0
                bool keeps_silence = true;
                if (n.has_curve && !n.table.empty()) {
                    const size_t cn = n.table.size();
                    keeps_silence = cn % 2 == 1 ? std::fabs(n.table[cn / 2]) < 1e-9f : std::fabs((n.table[cn / 2 - 1] + n.table[cn / 2]) / 2.f) < 1e-9f;
                }
                // a silent input that still produces sound does so on the ONE channel a silent quantum has (waveshaper.rs:395-400)
                auto shaper_lay = [&](const Lay& l) {
                    if (keeps_silence || !l.may_silent) return l;
                    return Lay{1, l.hi, 1, l.nhi, false};
                };
                if (n.oversample && n.has_curve) {  // waveshaper.rs:409-480: up-sample, shape, down-sample
                    // input that can fall silent: a curve that maps 0 to 0 makes the node return early WITHOUT feeding its resamplers
                    // (frozen state: the kernel then works on the last processed quanta); a curve that does not keeps processing — on the
                    // one channel of a silent quantum, which rebuilds the resamplers of a wider node (waveshaper.rs:395-420)
                    const bool freeze = in0.dyn() && keeps_silence && in0.nlo == in0.nhi && in0.nhi == ch;
                    const bool as_static = !in0.dyn() || (!keeps_silence && ch == 1 && in0.hi == 1);
                    if (!freeze && !as_static)
                        return bail(WAE_UNSUPPORTED, "an over-sampled WaveShaperNode whose input changes its channel count is not lowered to the GPU "
                                                     "(the reference rebuilds its resamplers then, waveshaper.rs:409-420)");
                    if (!need_out(ch)) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                    if (freeze) out_like_input();
                    const int factor = n.oversample == WAE_OVERSAMPLE_X2 ? 2 : 4;
                    auto& filt = os_filters[factor];
                    if (!filt.first) {
                        const std::vector<float2> fu = resampler_filter_bins(128, 128 * factor, 128);
                        const std::vector<float2> fd = resampler_filter_bins(128 * factor, 128, 128);
                        filt.first = upload(fu);
                        filt.second = upload(fd);
                        if (!dry) cudaStreamSynchronize(eng->stream);  // the host vectors go out of scope
                    }
                    ShaperOsInst so{};
                    so.in = p.in_buf[0];
                    so.out = p.out_buf[0];
                    so.curve = curve;
                    so.n = (int)n.table.size();
                    so.ch = ch;
                    so.factor = factor;
                    so.f_up = filt.first;
                    so.f_dn = filt.second;
                    so.hist = alloc<float>((size_t)256 * ch, true, true);
                    if (!so.hist || !so.f_up || !so.f_dn) return bail(WAE_OUT_OF_MEMORY, "out of device memory (over-sampled shaper)");
                    if (freeze) {
                        so.prev = alloc<int32_t>((size_t)(2 * (b->chunk / 128) + 2));
                        if (!so.prev) return bail(WAE_OUT_OF_MEMORY, "out of device memory (over-sampled shaper)");
                    }
                    StageBuild& os = stage(L, S_SHAPER_OS);
                    os.max_ch = std::max(os.max_ch, ch);
                    os.shaper_os.push_back(so);
                    break;
                }
                if (fuse_n) {
                    PendingChain pc = open_chain();
                    pc.inst.has_shaper = 1;
                    pc.inst.curve = curve;
                    pc.inst.shaper_n = (int)n.table.size();
                    pc.inst.shaper_keeps_silence = keeps_silence ? 1 : 0;
                    pc.lay = shaper_lay(pc.lay);
                    pc.phase = 5;
                    if (!finish_chain(std::move(pc))) return false;
                    break;
                }
                if (!need_out(ch)) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                ShaperInst sh{};
                if (in0.dyn() && !keeps_silence && n.has_curve) {
                    out_dynamic(shaper_lay(in0));
                    if (p.out_buf[0].meta) meta_stage(L, META_SHAPER, p.in_buf[0], ch, p.out_buf[0], ch);
                } else {
                    out_like_input();
                }
                sh.in = p.in_buf[0];
                sh.out = p.out_buf[0];
                sh.ch = ch;
                sh.n = (int)n.table.size();
                sh.curve = curve;
                stage(L, S_SHAPER).shaper.push_back(sh);
                break;
            }
            case K_SPANNER: {
                PRef ppan = param_ref(g, n.params[0]);
                float pan = ppan.v;
                int ch = p.in_ch[0];
                if (!need_out(2)) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                // silent in -> silent out, else two channels (stereo_panner.rs:230-235); the kernel picks the mono / stereo law per quantum
                out_dynamic(in0.may_silent ? Lay{1, 2, 2, 2, true} : Lay::fixed(2));
                if (p.out_buf[0].meta) meta_stage(L, META_PAN, p.in_buf[0], ch, p.out_buf[0], 2);
                float x = ch == 1 ? (pan + 1.f) * 0.5f : (pan <= 0.f ? pan + 1.f : pan);  // stereo_panner.rs:247-249,274-276
                float gl, gr;
                hm::stereo_gains(x, gl, gr);
                StageBuild& s = stage(L, S_SPAN);
                s.span.push_back(SPanInst{p.in_buf[0], p.out_buf[0], pan, ch, ppan.dyn ? ppan.track : BufRef{nullptr, 0, 0}});
                s.span_gains.push_back(make_float2(gl, gr));
                break;
            }
            case K_PANNER: {
                // the 15 spatial params (panner.rs:714-780): 6 of the node, 9 of the AudioListener (graph ids 2..10)
                PRef pr[15];
                bool moving = false;
                for (int i = 0; i < 15; i++) {
                    pr[i] = param_ref(g, i < 6 ? n.params[i] : (uint32_t)(2 + i - 6));
                    moving = moving || pr[i].dyn;
                }
                int ch = p.in_ch[0];
                // (a static HRTF panner with a constant-layout input is lowered to the convolver kernels, which take their own output buffer)
                static const bool hrtf_fft_on = [] { const char* e = getenv("WAE_HRTF_FFT"); return !e || atoi(e) != 0; }();
                const bool hrtf_as_conv = n.panning_model == WAE_PANNING_HRTF && hrtf_fft_on && eng->sphere && !moving && !in0.dyn() && cur_cls != 1 &&
                                          seg_start == 0 && seg_end >= b->lq && (ch == 1 || ch == 2);
                if (!hrtf_as_conv && !need_out(2)) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                if (hrtf_as_conv) p.out_lay = {Lay::fixed(2)};
                else out_dynamic(in0.may_silent ? Lay{1, 2, 2, 2, true} : Lay::fixed(2));  // panner.rs:698-708
                // (the HRTF panner keeps its own tail budget: its layout track is written by k_hrtf_map)
                if (!hrtf_as_conv && p.out_buf[0].meta && n.panning_model != WAE_PANNING_HRTF) meta_stage(L, META_PAN, p.in_buf[0], ch, p.out_buf[0], 2);
                spatial::PanModel model{};
                model.distance_model = n.distance_model;
                model.ref_distance = n.ref_distance;
                model.max_distance = n.max_distance;
                model.rolloff_factor = n.rolloff_factor;
                model.cone_inner_angle = n.cone_inner_angle;
                model.cone_outer_angle = n.cone_outer_angle;
                model.cone_outer_gain = n.cone_outer_gain;
                SpatialTracks tr{};
                float v[15];
                for (int i = 0; i < 15; i++) {
                    v[i] = tr.value[i] = pr[i].v;
                    tr.track[i] = pr[i].dyn ? pr[i].track : BufRef{nullptr, 0, 0};
                }
                const spatial::SpatialParams sp0 = spatial::spatial_params(model, v);  // static source and listener
                if (n.panning_model == WAE_PANNING_HRTF) {  // panner.rs:781-830
                    const HrirSphere* sph = eng->sphere;
                    if (!sph) return bail(WAE_UNSUPPORTED, "HRTF panning needs an HRIR sphere: call wae_engine_set_hrir_sphere first");
                    uint32_t sr = (uint32_t)g->sample_rate;
                    if (sr < 27000) sr = 27000;  // panner.rs:46
                    uint32_t taps = sph->taps;
                    const float* d_ir = eng->d_sphere_ir;
                    const float* h_ir = sph->ir.data();  // [vertex][2][taps] on the host
                    if (sr != sph->sample_rate) {  // the crate resamples the responses to the context rate once (wae_hrtf_host.h)
                        std::lock_guard<std::mutex> slk(eng->sphere_mu);
                        auto it = eng->sphere_rates.find(sr);
                        if (it == eng->sphere_rates.end()) {
                            const HrirSphere rs = sph->at_rate(sr);
                            wae_engine::RateSphere r;
                            r.taps = rs.taps;
                            if (r.taps < 2) return bail(WAE_UNSUPPORTED, "HRTF panning: the HRIR sphere is too short to be resampled to the context rate");
                            if (cudaMalloc(&r.d_ir, rs.ir.size() * sizeof(float)) != cudaSuccess) return bail(WAE_OUT_OF_MEMORY, "out of device memory (resampled HRIR sphere)");
                            cudaMemcpy(r.d_ir, rs.ir.data(), rs.ir.size() * sizeof(float), cudaMemcpyHostToDevice);
                            r.ir_host = rs.ir;
                            it = eng->sphere_rates.emplace(sr, r).first;
                        }
                        taps = it->second.taps;
                        d_ir = it->second.d_ir;
                        h_ir = it->second.ir_host.data();  // (map nodes are stable; entries are only dropped with the sphere)
                    }
                    // A static source heard by a static listener through a constant-layout input is ONE fixed pair of impulse responses:
                    // out_ear = gain * (h_ear * mono(in)).  The crate evaluates that by FFT overlap-save per 128-frame block (hrtf 0.8.1
                    // process_samples); here it is handed to the time-batched convolver kernels as a ConvolverNode-shaped problem —
                    // response = the blended pair with the gain (and the reference's correction of 2 for a two-channel input,
                    // panner.rs:805-812) folded in; a two-channel input is mixed down to mono by the forward transform's loads; one
                    // partition, so the product is formed inside the inverse transform — instead of 2 x taps multiply-adds per output
                    // frame in k_hrtf_fir.  WAE_HRTF_FFT=0: keep the FIR kernel.
                    if (hrtf_as_conv) {
                        float proj[3];
                        spatial::projected_source(sp0, proj);
                        const float dir[3] = {proj[0], proj[2], proj[1]};  // HrtfState::process swaps y / z (panner.rs:248-252)
                        HrtfSel sel{{0, 0, 0}, {0.f, 0.f, 0.f}, sp0.cone_gain * sp0.dist_gain, 0.f};
                        sph->locate(dir, sel.v, sel.w);  // no face: all-zero weights (silence)
                        PcmBuffer resp;
                        if (!resp.allocate(2, taps, false)) return bail(WAE_OUT_OF_MEMORY, "out of host memory (hrtf response)");
                        const float corr = ch == 2 ? 2.f : 1.f;  // overall_gain_correction of a two-channel input (panner.rs:805-812)
                        const float* A = h_ir + (size_t)sel.v[0] * 2 * taps;
                        const float* B = h_ir + (size_t)sel.v[1] * 2 * taps;
                        const float* C = h_ir + (size_t)sel.v[2] * 2 * taps;
                        for (uint32_t k = 0; k < taps; k++) {  // (the blend k_hrtf_fir does, same f32 operations)
                            const float l = (A[k] * sel.w[0] + B[k] * sel.w[1]) + C[k] * sel.w[2];
                            const float r = (A[taps + k] * sel.w[0] + B[taps + k] * sel.w[1]) + C[taps + k] * sel.w[2];
                            resp.channels[0].p[k] = corr * (l * sel.gain);
                            resp.channels[1].p[k] = corr * (r * sel.gain);
                        }
                        resp.sample_rate = (float)sr;
                        if (!plan_convolver(g, p, L, nullptr, -1, &resp)) return false;
                        break;
                    }
                    HrtfInst h{};
                    h.in = p.in_buf[0];
                    h.out = p.out_buf[0];
                    h.in_ch = ch;
                    h.L = (int)taps;
                    h.sphere_ir = d_ir;
                    h.sel = nullptr;
                    h.correction = ch == 2 ? 2.f : 1.f;
                    h.hist = alloc<float>(taps, true, true);
                    if (!h.hist) return bail(WAE_OUT_OF_MEMORY, "out of device memory (hrtf history)");
                    if (in0.dyn()) {  // the node stops processing (and freezes) once its tail budget is used up: panner.rs:697-711
                        h.dyn = 1;
                        h.cmap = alloc<int32_t>((size_t)(b->chunk / 128 + 2));
                        h.tail = alloc<int64_t>(1, true, true);
                        if (!h.cmap || !h.tail) return bail(WAE_OUT_OF_MEMORY, "out of device memory (hrtf layout)");
                    }
                    StageBuild& hs = stage(L, S_HRTF);
                    if (moving) {
                        HrtfSelInst si{};
                        si.sp = tr;
                        si.model = model;
                        si.pos = eng->d_sphere_pos;
                        si.tri = eng->d_sphere_tri;
                        si.n_faces = (int)(sph->tri.size() / 3);
                        si.sel = alloc<HrtfSel>((size_t)(b->chunk / 128 + 1));
                        if (!si.sel) return bail(WAE_OUT_OF_MEMORY, "out of device memory (hrtf selection)");
                        h.sel = si.sel;
                        hs.hrtf_sel.push_back(si);
                    } else {
                        float proj[3];
                        spatial::projected_source(sp0, proj);
                        const float dir[3] = {proj[0], proj[2], proj[1]};  // HrtfState::process swaps y / z (panner.rs:248-252)
                        h.static_sel = HrtfSel{{0, 0, 0}, {0.f, 0.f, 0.f}, sp0.cone_gain * sp0.dist_gain, 0.f};
                        sph->locate(dir, h.static_sel.v, h.static_sel.w);  // no face: all-zero weights (silence)
                    }
                    hs.hrtf.push_back(h);
                    break;
                }
                if (moving) {
                    PanDynInst d{};
                    d.in = p.in_buf[0];
                    d.out = p.out_buf[0];
                    d.sp = tr;
                    d.model = model;
                    d.in_ch = ch;
                    stage(L, S_PAN_DYN).pan_dyn.push_back(d);
                    break;
                }
                PanInst pi{};
                pi.in = p.in_buf[0];
                pi.out = p.out_buf[0];
                pi.in_ch = ch;
                pi.azimuth = sp0.azimuth;
                pi.dist_gain = sp0.dist_gain;
                pi.cone_gain = sp0.cone_gain;
                stage(L, S_PAN).pan.push_back(pi);
                break;
            }
            case K_DELAY_W: {
                p.out_ch = {p.in_ch[0]};
                p.out_buf = {p.in_buf[0]};
                p.out_lay = {in0};
                if (Orderer::contains(ord.broken, id)) {
                    // cycle breaker applied (graph.rs:458-466): the hidden writer->reader edge is gone, the reader ran
                    // earlier in this quantum from the ring; record this quantum's input now
                    delay_ch_seen[{gi, n.delay_peer}] = p.in_ch[0];
                    auto it = delay_rings.find({gi, id});
                    if (it == delay_rings.end()) return bail(WAE_UNSUPPORTED, "DelayNode writer processed before its reader inside a cycle");
                    if (it->second.ch != p.in_ch[0]) {
                        if (!dry) return bail(WAE_UNSUPPORTED, "channel layout of a DelayNode in a feedback cycle did not converge");
                        break;  // sizing pass: the hint is corrected and the pass repeated
                    }
                    DelayInst d{};
                    d.in = p.in_buf[0];
                    d.ch = it->second.ch;
                    d.ring = it->second.ring;
                    d.ring_len = it->second.ring_len;
                    d.mono_at = it->second.mono_at;
                    d.mono_len = it->second.mono_len;
                    d.dyn = it->second.mono_at ? 3 : 0;  // inside a cycle the writer runs after the reader: it extends the one-channel track
                    stage(L, S_DELAY_WRITE).delay.push_back(d);
                }
                break;
            }
            case K_DELAY_R: {
                PRef pdl = param_ref(g, n.params[0]);
                float dt = pdl.v;
                const bool in_cycle = Orderer::contains(ord.broken, n.delay_peer);
                int ch = p.in_ch[0];
                if (in_cycle) {
                    ch = 1;
                    if (delay_ch_hint) {
                        auto it = delay_ch_hint->find({gi, id});
                        if (it != delay_ch_hint->end()) ch = it->second;
                    }
                }
                if (!need_out(ch)) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                DelayInst d{};
                d.in = p.in_buf[0];
                d.out = p.out_buf[0];
                d.ch = ch;
                d.in_cycle = in_cycle ? 1 : 0;
                if (pdl.dyn) d.delay_track = pdl.track;
                d.sample_rate = g->sample_rate;
                double delay = (double)dt;
                if (in_cycle) delay = std::max(delay, 128. / sr);  // delay.rs:699-703: at least one quantum inside a cycle
                double num_samples = delay * sr;               // delay.rs:706
                double position = 0. - num_samples;            // sample_index 0
                double pf = std::floor(position);
                d.fl = (int64_t)pf;
                d.k = (float)(position - pf);
                uint64_t max_frames = (uint64_t)std::ceil(std::max(n.max_delay_time, 128. / sr) * sr) + 2;
                d.ring_len = next_pow2(max_frames + 128);
                d.ring = alloc<float>((size_t)ch * d.ring_len, true, true);
                if (!d.ring) return bail(WAE_OUT_OF_MEMORY, "out of device memory (delay ring)");
                b->arena_bytes += (size_t)ch * d.ring_len * 4;
                if (ch <= 2) {
                    // The reader reports a quantum without any normal sample as silent (delay.rs:654-664) and the ring follows the
                    // channel count of the writer's input (:470-488): its output layout is never constant.  (Wider than stereo: the
                    // static layout is kept, the re-mix of the ring is not followed.)
                    d.dyn = 1;
                    d.mono_len = (int32_t)next_pow2((uint64_t)(b->chunk / 128 + 2));
                    d.mono_at = alloc<int64_t>((size_t)d.mono_len, true, true);
                    if (!d.mono_at) return bail(WAE_OUT_OF_MEMORY, "out of device memory (delay layout track)");
                    const Lay wl = in_cycle ? Lay{1, (uint8_t)ch, 1, (uint8_t)ch, true} : in0;
                    out_dynamic(Lay{1, (uint8_t)ch, (uint8_t)(wl.dyn() ? 1 : ch), (uint8_t)ch, true});
                    d.out = p.out_buf[0];
                    if (!in_cycle) stage(L, S_DELAY_MONO).delay.push_back(d);
                }
                stage(L, S_DELAY).delay.push_back(d);
                if (in_cycle) delay_rings[{gi, n.delay_peer}] = DelayRing{d.ring, d.ring_len, ch, d.mono_at, d.mono_len};
                else stage(L, S_DELAY_WRITE).delay.push_back(d);  // acyclic: history is recorded right after the read
                break;
            }
            case K_COMP: {
                PRef cp[5];
                for (int i = 0; i < 5; i++) cp[i] = param_ref(g, n.params[i]);
                const float at = cp[0].v, kn = cp[1].v, ra = cp[2].v, re = cp[3].v, th = cp[4].v;
                int ch = p.in_ch[0];
                if (!need_out(ch)) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                CompInst c{};
                c.in = p.in_buf[0];
                c.out = p.out_buf[0];
                c.ch = ch;
                int ring_size = (int)std::ceil(g->sample_rate * 0.006f / 128.f) + 1;  // dynamics_compressor.rs:250-255
                c.delay_frames = (ring_size - 1) * 128;
                c.ring_len = next_pow2((uint64_t)c.delay_frames + 128);
                c.ring = alloc<float>((size_t)ch * c.ring_len, true, true);
                c.state = alloc<float>(2, true, true);
                c.meta_ring = alloc<uint8_t>(8, true, true);
                if (!c.ring || !c.state || !c.meta_ring) return bail(WAE_OUT_OF_MEMORY, "out of device memory (compressor)");
                // the look-ahead ring starts out silent and hands on the layout of the quantum it delays (dynamics_compressor.rs:340-349,452-468)
                out_dynamic(Lay{1, in0.hi, in0.nlo, in0.nhi, true});
                c.out = p.out_buf[0];
                c.threshold = th; c.knee = kn; c.ratio = ra; c.attack = at; c.release = re;
                for (int i = 0; i < 5; i++) c.track[i] = cp[i].dyn ? cp[i].track : BufRef{nullptr, 0, 0};
                c.sample_rate = g->sample_rate;
                stage(L, S_COMP).comp.push_back(c);
                if (!dry) {
                    std::lock_guard<std::recursive_mutex> lk(b->mu);  // (groups are planned on worker threads)
                    bool known = false;
                    for (auto& r : b->compressors) known = known || (r.graph == gi && r.node == id);
                    if (!known) b->compressors.push_back(wae_batch::CompRec{gi, id, c.state});
                }
                break;
            }
            case K_ANALYSER: {
                int ch = p.in_ch[0];
                // pass-through (analyser.rs:267-294): the output IS the input buffer (nobody writes an edge buffer after its
                // producer), only the ring is written
                p.out_ch = {ch};
                p.out_buf = {p.in_buf[0]};
                p.out_lay = {in0};
                AnalyserInst a{};
                a.in = p.in_buf[0];
                a.out = BufRef{nullptr, 0, 0};
                a.ch = ch;
                a.ring = alloc<float>(32768 + 128, true, true);
                if (!a.ring) return bail(WAE_OUT_OF_MEMORY, "out of device memory (analyser ring)");
                stage(L, S_ANALYSER).analyser.push_back(a);
                {
                    float* last = alloc<float>(16384, true, true);
                    float* db = alloc<float>(16384);
                    if (!last || !db) return bail(WAE_OUT_OF_MEMORY, "out of device memory (analyser)");
                    if (!dry) {
                        std::lock_guard<std::recursive_mutex> lk(b->mu);
                        bool known = false;
                        for (auto& r : b->analysers) known = known || (r.graph_index == gi && r.node == id);
                        if (!known) b->analysers.push_back(AnalyserRec{gi, id, a.ring, n.fft_size, n.smoothing, last, db, false, n.min_db, n.max_db});
                    }
                }
                algorithmic_bytes += (uint64_t)b->lq * 4;  // ring write, SURVEY §8(d)
                break;
            }
            case K_MERGER: {
                int k = n.n_inputs;
                if (!need_out(k)) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                {
                    // `k` channels as soon as one input is not silent, else silent (channel_merger.rs:160-168)
                    bool some_always_on = false, any_dyn = false;
                    for (int i = 0; i < k; i++) {
                        some_always_on = some_always_on || !p.in_lay[i].may_silent;
                        any_dyn = any_dyn || p.in_lay[i].dyn();
                    }
                    if (any_dyn && !some_always_on) {
                        out_dynamic(Lay{1, (uint8_t)k, (uint8_t)k, (uint8_t)k, true});
                        if (p.out_buf[0].meta) {
                            MetaInst m{};
                            m.out = p.out_buf[0];
                            m.mode = META_MERGE;
                            m.out_ch = k;
                            m.count = k;
                            m.n_more = k;
                            m.more = upload(p.in_buf);
                            if (!m.more) return bail(WAE_OUT_OF_MEMORY, "out of device memory (merger inputs)");
                            stage(L, S_META).meta.push_back(m);
                        }
                    }
                }
                for (int i = 0; i < k; i++) stage(L, S_ROUTE).route.push_back(RouteInst{p.in_buf[i], p.out_buf[0], 0, i, 0, 1});
                break;
            }
            case K_SPLITTER: {
                int k = n.n_outputs;
                p.out_ch.assign(k, 1);
                p.out_buf.resize(k);
                p.out_lay.assign(k, Lay::fixed(1));
                for (int i = 0; i < k; i++) {
                    if (i < p.in_ch[0] && in0.dyn()) {  // channel i exists only in some quanta: copy it, zeros elsewhere, own layout track
                        p.out_buf[i] = arena_buf(1, true);
                        if (!p.out_buf[i].p) return bail(WAE_OUT_OF_MEMORY, "out of device memory (arena)");
                        p.out_lay[i] = Lay{1, 1, 1, 1, true};
                        stage(L, S_ROUTE).route.push_back(RouteInst{p.in_buf[0], p.out_buf[i], i, 0, 0, p.in_ch[0]});
                        meta_stage(L, META_SPLIT, p.in_buf[0], p.in_ch[0], p.out_buf[i], 1, 0, i);
                    } else if (i < p.in_ch[0]) {  // alias channel i of the input
                        BufRef r = p.in_buf[0];
                        r.p += (size_t)i * r.stride;
                        p.out_buf[i] = r;
                    } else {
                        p.out_buf[i] = arena_buf(1);
                        stage(L, S_ROUTE).route.push_back(RouteInst{p.in_buf[0], p.out_buf[i], 0, 0, 1, 0});
                    }
                }
                break;
            }
            case K_CONV: {
                // the destination's only input (and this node's only consumer): the inverse transforms write the rendered PCM
                const BufRef* dest = nullptr;
                BufRef fin{b->d_out + (size_t)gi * b->channels * b->length, (uint32_t)b->length, 1};
                if (fuse && cur_cls == 0 && b->length <= 0xffffffffull) {
                    int n_out = 0;
                    uint32_t to = 0;
                    int to_port = -1;
                    for (auto& e : ord.edges.at(id))
                        if (e.other_index >= 0) n_out++, to = e.other_id, to_port = e.other_index;
                    if (n_out == 1 && to_port == 0 && g->nodes.at(to).kind == K_DEST && pn.at(to).in_edges[0].size() == 1 &&
                        computed_channels(g->nodes.at(to).cfg, (n.buffer && n.buffer->channels.size() == 1 && p.in_ch[0] == 1) ? 1 : 2) == (int)b->channels)
                        dest = &fin;
                }
                if (!plan_convolver(g, p, L, dest, (int64_t)b->length)) return false;
                break;
            }
            default: return bail(WAE_UNSUPPORTED, "node kind not lowered to the GPU");
        }
    }
    // chains whose single consumer never showed up as an audio input (e.g. it feeds an AudioParam): materialise
    while (!pending.empty())
        if (!materialize(pending.begin()->first)) return false;
    return true;
}

template <typename T>
static void* up(wae_batch* b, const std::vector<T>& v) {
    return (void*)b->dupload(v);
}

// While alive, `g->nodes` is the graph description that is valid at `frame` (the snapshot taken at the first suspend point
// after it, or the live graph when none follows): the planner reads g->nodes without knowing about suspend points.
struct EpochView {
    wae_graph* g;
    size_t e;
    EpochView(wae_graph* g_, int64_t frame) : g(g_), e(0) {
        while (e < g->epochs.size() && (int64_t)g->epochs[e].frame <= frame) e++;
        if (e < g->epochs.size()) std::swap(g->nodes, g->epochs[e].nodes);
    }
    ~EpochView() {
        if (e < g->epochs.size()) std::swap(g->nodes, g->epochs[e].nodes);
    }
    EpochView(const EpochView&) = delete;
    EpochView& operator=(const EpochView&) = delete;
};

// WAE_OPT_BIND_NUMA / WAE_BIND_NUMA=1: restrict the calling thread (and the engine's workers, which inherit the mask) to the CPUs
// next to this engine's GPU (/sys/bus/pci/devices/<bus id>/local_cpulist), so that page-locked staging memory is allocated and
// touched on that socket: on a two-socket host the H2D / D2H copies of eight ranks otherwise share one socket's memory
// controllers and the inter-socket link.
static bool bind_to_device_numa_node(wae_engine* eng) {
    char bus[64] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, eng->device) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    for (char* c = bus; *c; c++) *c = (char)std::tolower((unsigned char)*c);
    std::ifstream f(std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist");
    std::string list;
    if (!f || !std::getline(f, list) || list.empty()) return false;
    cpu_set_t set;
    CPU_ZERO(&set);
    std::stringstream ss(list);
    std::string part;
    int n = 0;
    while (std::getline(ss, part, ',')) {
        int a = 0, b2 = 0;
        if (std::sscanf(part.c_str(), "%d-%d", &a, &b2) == 2) {
            for (int c = a; c <= b2 && c < CPU_SETSIZE; c++, n++) CPU_SET(c, &set);
        } else if (std::sscanf(part.c_str(), "%d", &a) == 1 && a < CPU_SETSIZE) {
            CPU_SET(a, &set);
            n++;
        }
    }
    if (n == 0 || sched_setaffinity(0, sizeof set, &set) != 0) return false;
    eng->numa_cpus = list;
    return true;
}

}  // namespace

extern "C" {

WAE_API wae_status wae_engine_create(int32_t device_ordinal, wae_engine** out) {
    if (!out) return fail(WAE_INVALID_ARGUMENT, "null out pointer");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(WAE_NO_DEVICE, std::string("no CUDA device usable (") + cudaGetErrorString(e) + "): this library has no CPU fallback");
    if (device_ordinal < 0 || device_ordinal >= count) return fail(WAE_NO_DEVICE, "device ordinal out of range");
    CUDA_TRY(cudaSetDevice(device_ordinal));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device_ordinal));
    if (prop.major < 10) return fail(WAE_NO_DEVICE, "the kernels are built for sm_100a only");
    auto* eng = new wae_engine;
    eng->device = device_ordinal;
    CUDA_TRY(cudaStreamCreateWithFlags(&eng->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&eng->s_h2d, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&eng->s_d2h, cudaStreamNonBlocking));
    if (const char* e = getenv("WAE_BIND_NUMA"))
        if (atoi(e) != 0) bind_to_device_numa_node(eng);
    std::vector<float> sine = hm::sine_table();
    CUDA_TRY(cudaMalloc(&eng->d_sine, sine.size() * sizeof(float)));
    CUDA_TRY(cudaMemcpy(eng->d_sine, sine.data(), sine.size() * sizeof(float), cudaMemcpyHostToDevice));
    upload_twiddles();
    CUDA_TRY(cudaDeviceSynchronize());
    *out = eng;
    return WAE_OK;
}

WAE_API wae_status wae_engine_destroy(wae_engine* eng) {
    if (!eng) return WAE_OK;
    cudaSetDevice(eng->device);
    if (eng->d_sine) cudaFree(eng->d_sine);
    if (eng->d_sphere_ir) cudaFree(eng->d_sphere_ir);
    if (eng->d_sphere_pos) cudaFree(eng->d_sphere_pos);
    if (eng->d_sphere_tri) cudaFree(eng->d_sphere_tri);
    eng->drop_rate_spheres();
    delete eng->sphere;
    delete eng->pool;  // joins the workers
    eng->dev_trim();
    for (int i = 0; i < wae_engine::kStageSlots; i++)
        if (eng->h_stage[i]) cudaFreeHost(eng->h_stage[i]);
    if (eng->h_ring) cudaFreeHost(eng->h_ring);
    if (eng->stream) cudaStreamDestroy(eng->stream);
    if (eng->s_h2d) cudaStreamDestroy(eng->s_h2d);
    if (eng->s_d2h) cudaStreamDestroy(eng->s_d2h);
    delete eng;
    return WAE_OK;
}

// load_hrtf_processor (src/node/panner.rs:39-68): the reference embeds resources/IRC_1003_C.bin; the binding hands the same
// bytes to the engine once.  Batches prepared afterwards may contain PanningModelType::HRTF panners.
WAE_API wae_status wae_engine_set_hrir_sphere(wae_engine* eng, const void* data, uint64_t len) {
    if (!eng) return fail(WAE_INVALID_ARGUMENT, "null engine");
    auto* sp = new HrirSphere();
    std::string err;
    if (!sp->parse(static_cast<const uint8_t*>(data), len, err)) {
        delete sp;
        return fail(WAE_INVALID_ARGUMENT, err.c_str());
    }
    CUDA_TRY(cudaSetDevice(eng->device));
    float* d = nullptr;
    if (cudaMalloc(&d, sp->ir.size() * sizeof(float)) != cudaSuccess) {
        delete sp;
        return fail(WAE_OUT_OF_MEMORY, "out of device memory (HRIR sphere)");
    }
    float* dpos = nullptr;
    uint32_t* dtri = nullptr;
    if (cudaMalloc(&dpos, sp->pos.size() * sizeof(float)) != cudaSuccess || cudaMalloc(&dtri, sp->tri.size() * sizeof(uint32_t)) != cudaSuccess) {
        cudaFree(d);
        if (dpos) cudaFree(dpos);
        delete sp;
        return fail(WAE_OUT_OF_MEMORY, "out of device memory (HRIR sphere)");
    }
    cudaMemcpy(d, sp->ir.data(), sp->ir.size() * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(dpos, sp->pos.data(), sp->pos.size() * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(dtri, sp->tri.data(), sp->tri.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
    CUDA_TRY(cudaStreamSynchronize(eng->stream));  // batches in flight may still read the previous sphere
    if (eng->d_sphere_ir) cudaFree(eng->d_sphere_ir);
    if (eng->d_sphere_pos) cudaFree(eng->d_sphere_pos);
    if (eng->d_sphere_tri) cudaFree(eng->d_sphere_tri);
    eng->drop_rate_spheres();
    eng->d_sphere_pos = dpos;
    eng->d_sphere_tri = dtri;
    delete eng->sphere;
    eng->sphere = sp;
    eng->d_sphere_ir = d;
    return WAE_OK;
}

// the CUDA stream every kernel of this engine is launched on (for callers that time with their own events)
WAE_API wae_status wae_engine_stream(wae_engine* eng, void** out_stream) {
    *out_stream = (void*)eng->stream;
    return WAE_OK;
}

WAE_API wae_status wae_engine_set_option(wae_engine* eng, uint32_t option, int64_t value) {
    switch (option) {
        case WAE_OPT_CHUNK_FRAMES:
            if (value < 0 || value % 128 != 0) return fail(WAE_INVALID_ARGUMENT, "chunk frames must be a multiple of 128");
            eng->chunk_frames = value;
            return WAE_OK;
        case WAE_OPT_FUSE: eng->fuse = value != 0; return WAE_OK;
        case WAE_OPT_VOICE_SUM: eng->voice_sum = value < 0 ? 0 : (value > 2 ? 2 : (int)value); return WAE_OK;
        case WAE_OPT_SERIAL_FILTERS: eng->serial_filters = value != 0; return WAE_OK;
        case WAE_OPT_PIPELINE_GROUPS:
            if (value < 0 || value > 1024) return fail(WAE_INVALID_ARGUMENT, "pipeline groups must be in [0, 1024]");
            eng->pipeline_groups = (int)value;
            return WAE_OK;
        case WAE_OPT_PARAM_PARALLEL: eng->param_parallel = value > 2 ? 2 : (int)value; return WAE_OK;
        case WAE_OPT_BIND_NUMA:
            if (value != 0 && eng->pool) return fail(WAE_INVALID_STATE, "bind the engine to its NUMA node before its first render (worker threads already run)");
            if (value != 0 && !bind_to_device_numa_node(eng)) return fail(WAE_UNSUPPORTED, "could not read / apply the CPU list of the GPU's NUMA node");
            return WAE_OK;
        case WAE_OPT_HOST_WORKERS:
            if (value < 0 || value > 256) return fail(WAE_INVALID_ARGUMENT, "host workers must be in [0, 256]");
            if (eng->pool) return fail(WAE_INVALID_STATE, "the worker threads already run");
            eng->n_workers = (int)value;
            return WAE_OK;
        case WAE_OPT_CHAIN_TMA: chain_set_tuning(value != 0 ? 1 : 0, -1); return WAE_OK;
        case WAE_OPT_CHAIN_WAVES:
            if (value < 0 || value > 1024) return fail(WAE_INVALID_ARGUMENT, "chain waves must be in [0, 1024]");
            chain_set_tuning(-1, (int)value);
            return WAE_OK;
        case WAE_OPT_CHAIN_PREPASS: chain_set_prepass(value != 0 ? 1 : 0); return WAE_OK;
        default: return fail(WAE_INVALID_ARGUMENT, "unknown option");
    }
}

WAE_API wae_status wae_batch_destroy(wae_batch* b) {
    if (!b) return WAE_OK;
    cudaSetDevice(b->engine->device);
    if (b->engine->stream) {  // (a plan-only batch has no streams)
        cudaStreamSynchronize(b->engine->stream);
        cudaStreamSynchronize(b->engine->s_h2d);
        cudaStreamSynchronize(b->engine->s_d2h);
    }
    for (void* h : b->pinned) cudaFreeHost(h);
    for (auto& g : b->groups) {
        if (g.ev_h2d) cudaEventDestroy(g.ev_h2d);
        if (g.ev_done) cudaEventDestroy(g.ev_done);
    }
    for (void* p : b->allocs) b->engine->dev_release(p);  // kept by the engine for the next batch (wae_engine::dev_alloc)
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
    for (auto e : b->stage_events) cudaEventDestroy(e);
    delete b;
    return WAE_OK;
}

// Graph::order_nodes as the planner runs it (host only): lets the CPU tests pin the ordering against the reference's tests
WAE_API wae_status wae_graph_render_order(wae_graph* g, wae_node_id* ids, uint32_t cap, uint32_t* n) {
    if (!g || !n || (cap && !ids)) return fail(WAE_INVALID_ARGUMENT, "null argument");
    Orderer o;
    o.g = g;
    o.run();
    *n = (uint32_t)o.ordered.size();
    for (uint32_t i = 0; i < *n && i < cap; i++) ids[i] = o.ordered[i];
    return WAE_OK;
}

// ---- wae_batch_prepare in phases ------------------------------------------------------------------------------------------
// A (prep_begin): validation, graph groups, the sizing pass of the planner (no device memory touched; groups in parallel on the
//   engine's worker threads), chunk size.  B (prep_plan_group): the real plan of ONE group — allocation of its node state / arena,
//   upload of its instance tables — safe to run for several groups at once (shared structures are guarded by wae_batch::mu).
//   C (prep_append_group / prep_finish): the groups' stages joined in group order, statistics.
// wae_batch_prepare runs A, B for every group, C.  The one-shot render (wae_render_batch with a host buffer) runs A, starts the
// H2D copies of the source PCM, and then overlaps B of group k+1.. with the copies and the render of group k.
struct PrepState {
    std::map<std::pair<uint32_t, uint32_t>, int> delay_ch_hint;
    std::unordered_map<uint64_t, Planner::IrSpectra> ir_cache;
    uint64_t algorithmic_bytes = 0;
    std::chrono::steady_clock::time_point t0, t1;
};
struct GroupPlan {  // result of phase B for one group
    std::vector<Stage> stages;
    std::vector<std::pair<size_t, size_t>> seg_ranges;  // per segment: [first, last) into `stages`
    uint64_t algorithmic_bytes = 0;
    int code = WAE_OK;
    std::string error;
};

static wae_status prep_begin(wae_engine* eng, wae_graph* const* graphs, uint32_t n_graphs, wae_plan_info* plan, wae_batch** out_b, PrepState& ps,
                             bool want_d_out) {
    if (!eng || !graphs || n_graphs == 0) return fail(WAE_INVALID_ARGUMENT, "null / empty batch");
    if (!plan) CUDA_TRY(cudaSetDevice(eng->device));
    for (uint32_t i = 0; i < n_graphs; i++) {
        if (!graphs[i]) return fail(WAE_INVALID_ARGUMENT, "null graph in the batch");
        if (graphs[i]->channels != graphs[0]->channels || graphs[i]->length != graphs[0]->length ||
            graphs[i]->sample_rate != graphs[0]->sample_rate)
            return fail(WAE_INVALID_ARGUMENT, "all graphs of a batch must share number_of_channels, length and sample_rate");
    }
    ps.t0 = std::chrono::steady_clock::now();
    auto* b = new wae_batch;
    b->engine = eng;
    b->s_h2d = eng->s_h2d;
    b->s_d2h = eng->s_d2h;
    b->n_graphs = n_graphs;
    b->channels = graphs[0]->channels;
    b->length = graphs[0]->length;
    b->lq = (int64_t)((b->length + 127) / 128 * 128);
    bool has_conv = false, has_hrtf = false;
    std::vector<char> graph_has_conv(n_graphs, 0);
    std::vector<std::vector<int64_t>> cuts(n_graphs);  // per graph: its suspend frames inside the render
    for (uint32_t i = 0; i < n_graphs; i++) {
        for (auto& kv : graphs[i]->nodes) {
            if (kv.second.kind == K_CONV && kv.second.buffer) graph_has_conv[i] = 1;
            if (kv.second.kind == K_PANNER && kv.second.panning_model == WAE_PANNING_HRTF) has_hrtf = true;  // (static ones ride the convolver kernels)
        }
        for (auto& ep : graphs[i]->epochs) {
            if ((int64_t)ep.frame > 0 && (int64_t)ep.frame < b->lq) cuts[i].push_back((int64_t)ep.frame);
            for (auto& kv : ep.nodes)
                if (kv.second.kind == K_CONV && kv.second.buffer) graph_has_conv[i] = 1;
        }
        has_conv = has_conv || graph_has_conv[i];
    }
    size_t out_floats = (size_t)n_graphs * b->channels * b->length;
    if (!plan && want_d_out) {
        b->d_out = b->dalloc<float>(out_floats, true);
        if (!b->d_out) {
            wae_batch_destroy(b);
            return fail(WAE_OUT_OF_MEMORY, "out of device memory (output PCM)");
        }
    }
    // graph groups for the H2D / render / D2H pipeline
    int n_groups = eng->pipeline_groups;
    if (n_groups == 0) n_groups = n_graphs >= 512 ? 32 : (n_graphs >= 64 ? 8 : 1);  // measured on C2: 8 groups 88 ms, 16: 83.7, 32: 80.6 (fill / drain of the 3-stage pipeline)
    if (eng->pipeline_groups == 0 && n_graphs >= 512) {  // (tuning: WAE_AUTO_GROUPS overrides the automatic choice for large batches)
        static const int env_groups = [] { const char* e = getenv("WAE_AUTO_GROUPS"); return e ? atoi(e) : 0; }();
        if (env_groups > 0) n_groups = env_groups;
    }
    n_groups = std::max(1, std::min<int>(n_groups, (int)n_graphs));
    {
        // contiguous runs of graphs with the same suspend frames, cut further into about n_groups pieces
        const uint32_t target = (n_graphs + (uint32_t)n_groups - 1) / (uint32_t)n_groups;
        uint32_t g0 = 0;
        for (uint32_t i = 1; i <= n_graphs; i++)
            if (i == n_graphs || cuts[i] != cuts[g0] || i - g0 >= target) {
                wae_batch::Group grp;
                grp.g0 = g0;
                grp.g1 = i;
                grp.seg_bounds.push_back(0);
                for (int64_t f : cuts[g0]) grp.seg_bounds.push_back(f);
                grp.seg_bounds.push_back(b->lq);
                b->groups.push_back(grp);
                g0 = i;
            }
        n_groups = (int)b->groups.size();
    }
    // the convolver kernels work on whole partitions: a chunk (and so a render segment) has to start on one
    for (auto& grp : b->groups) {
        bool conv = false;
        for (uint32_t i = grp.g0; i < grp.g1; i++) conv = conv || graph_has_conv[i];
        if (!conv) continue;
        for (int64_t f : grp.seg_bounds)
            if (f != b->lq && f % WAE_CONV_BLOCK != 0) {
                wae_batch_destroy(b);
                return fail(WAE_UNSUPPORTED, "a suspend point that is not a multiple of the convolver partition (8192 frames) in a graph with a ConvolverNode is not lowered to the GPU");
            }
    }
    // Sizing pass (no device memory touched): arena floats per frame of the largest group and the source-PCM slab of
    // every group.  Chunk size: explicit option, else chosen so that a group's arena stays around 1 GiB.  A plan without any
    // arena buffer (fully fused source->...->destination chains) renders the whole length in one launch per group.
    // Convolvers work on whole 8192-frame blocks and prefer long chunks (their spectra ring, not the arena, is the traffic
    // that matters).
    b->chunk = 2048;
    uint64_t fpf = 0;
    bool has_feedback = false;
    std::vector<std::vector<int>> plan_stage_lists;  // plan-only: stage kinds per (group, segment) of the last iteration
    struct SizeOut {
        uint64_t fpf = 0;
        bool has_feedback = false;
        std::map<std::pair<uint32_t, uint32_t>, int> delay_ch_seen;
        std::vector<std::vector<int>> stage_lists;
        uint64_t digest = 1469598103934665603ull;
        int code = WAE_OK;
        std::string error;
    };
    WorkerPool* pool = (!plan && n_groups > 1 && n_graphs >= 64) ? eng->workers() : nullptr;
    bool converged = false;
    for (int iter = 0; iter < 8; iter++) {  // repeated only while the channel layout of in-cycle delays changes
        bool hints_changed = false;
        fpf = 0;
        plan_stage_lists.clear();
        std::vector<SizeOut> so(n_groups);
        // A group without suspend points whose graphs are large is sized by several workers, each with its own dry planner over a
        // contiguous run of the group's graphs: the graphs of a group share nothing in this pass but the running totals — arena floats per
        // frame (summed) and the cursor into the source-PCM slab (the copies are recorded relative to the run and rebased in graph order,
        // which is the order the planning pass walks).  WAE_PLAN_PARALLEL=1 lets wae_batch_plan (no engine, no GPU) size every group
        // both ways and compare — the check of this path in the CPU suite.
        struct RangeOut {
            uint64_t fpf = 0;
            size_t src_floats = 0;
            bool has_feedback = false;
            std::map<std::pair<uint32_t, uint32_t>, int> delay_ch_seen;
            std::vector<wae_batch::Group::SrcCopy> copies;
            std::vector<size_t> graph_base;  // slab cursor before each graph (relative to the run, rebased by the merge)
            Builds builds;                   // kept for the self-check only
            int code = WAE_OK;
            std::string error;
        };
        // `cursor0`: where the run starts in the slab — 0 while that is not known yet (the sizing pass proper: rebased by the merge), the
        // recorded base of graph i0 once it is (the planning pass and the self-check of the merged stage builds)
        auto size_range = [&](int k, uint32_t i0, uint32_t i1, RangeOut& ro, size_t cursor0) {  // single-segment groups only
            Planner sizing{b, eng};
            sizing.dry = true;
            sizing.group_graphs = (int)(b->groups[k].g1 - b->groups[k].g0);
            sizing.delay_ch_hint = &ps.delay_ch_hint;
            sizing.d_src = reinterpret_cast<float*>(uintptr_t(256));
            sizing.src_copies = &ro.copies;
            sizing.src_cursor = cursor0;
            sizing.begin_segment(0, b->lq);
            for (uint32_t i = i0; i < i1; i++) {
                EpochView view(graphs[i], 0);
                ro.graph_base.push_back(sizing.src_cursor);
                if (!sizing.plan_graph(graphs[i], i)) {
                    ro.code = sizing.error_code;
                    ro.error = sizing.error;
                    return;
                }
            }
            if (plan) ro.builds = std::move(sizing.builds);
            ro.fpf = sizing.arena_floats_per_frame;
            ro.src_floats = sizing.src_cursor - cursor0;
            ro.has_feedback = sizing.has_feedback;
            ro.delay_ch_seen = std::move(sizing.delay_ch_seen);
        };
        // (RangeOut of the whole group) from `parts` runs sized on `wp`; false: a run failed (first failing graph's error in `out`)
        auto size_group_split = [&](int k, WorkerPool* wp, int parts, RangeOut& out, const std::vector<size_t>* known_base = nullptr) {
            const uint32_t g0 = b->groups[k].g0, n = b->groups[k].g1 - g0;
            std::vector<RangeOut> ro((size_t)parts);
            auto run = [&](int t) {
                const uint32_t i0 = (uint32_t)((uint64_t)n * t / parts), i1 = (uint32_t)((uint64_t)n * (t + 1) / parts);
                size_range(k, g0 + i0, g0 + i1, ro[t], known_base ? (*known_base)[i0] : 0);
            };
            if (wp) wp->parallel_for(parts, run);
            else
                for (int t = 0; t < parts; t++) run(t);
            for (auto& r : ro) {
                if (r.code != WAE_OK) {
                    out.code = r.code;
                    out.error = r.error;
                    return false;
                }
                const size_t rebase = known_base ? 0 : out.src_floats;
                for (auto& c : r.copies) out.copies.push_back(wae_batch::Group::SrcCopy{c.buf, c.offset + rebase, c.floats});
                for (size_t gb : r.graph_base) out.graph_base.push_back(gb + rebase);
                if (plan) merge_builds(out.builds, r.builds);
                out.fpf += r.fpf;
                out.src_floats += r.src_floats;
                out.has_feedback = out.has_feedback || r.has_feedback;
                for (auto& kv : r.delay_ch_seen) out.delay_ch_seen[kv.first] = kv.second;
            }
            return true;
        };
        auto group_nodes = [&](int k) {
            size_t nn = 0;
            for (uint32_t i = b->groups[k].g0; i < b->groups[k].g1; i++) nn += graphs[i]->nodes.size();
            return nn;
        };
        static const bool check_split = [] { const char* e = getenv("WAE_PLAN_PARALLEL"); return e && atoi(e) != 0; }();
        auto size_group = [&](int k) {
            const uint32_t n_in_group = b->groups[k].g1 - b->groups[k].g0;
            const bool one_segment = b->groups[k].seg_bounds.size() == 2;
            // groups are sized one after the other on this thread (no group-level pool): its workers are free for the runs of a group
            if (!plan && !pool && one_segment && n_in_group >= 2 && group_nodes(k) >= 4096) {
                WorkerPool* wp = eng->workers();
                RangeOut out;
                if (!size_group_split(k, wp, (int)std::min<uint32_t>(n_in_group, (uint32_t)wp->size()), out)) {
                    so[k].code = out.code;
                    so[k].error = out.error;
                    return;
                }
                b->groups[k].src_copies = std::move(out.copies);
                b->groups[k].graph_src_base = std::move(out.graph_base);
                b->groups[k].src_floats = out.src_floats;
                so[k].fpf = out.fpf;
                so[k].has_feedback = out.has_feedback;
                so[k].delay_ch_seen = std::move(out.delay_ch_seen);
                return;
            }
            Planner sizing{b, eng};
            sizing.dry = true;
            sizing.group_graphs = (int)n_in_group;
            sizing.delay_ch_hint = &ps.delay_ch_hint;
            sizing.d_src = reinterpret_cast<float*>(uintptr_t(256));
            b->groups[k].src_copies.clear();
            b->groups[k].graph_src_base.clear();
            sizing.src_copies = &b->groups[k].src_copies;
            const std::vector<int64_t>& bounds = b->groups[k].seg_bounds;
            uint64_t serial_digest = 0;
            for (size_t sg = 0; sg + 1 < bounds.size(); sg++) {
                sizing.begin_segment(bounds[sg], bounds[sg + 1]);
                for (uint32_t i = b->groups[k].g0; i < b->groups[k].g1; i++) {
                    EpochView view(graphs[i], bounds[sg]);
                    if (one_segment) b->groups[k].graph_src_base.push_back(sizing.src_cursor);
                    if (!sizing.plan_graph(graphs[i], i)) {
                        so[k].code = sizing.error_code;
                        so[k].error = sizing.error;
                        return;
                    }
                }
                so[k].fpf = std::max(so[k].fpf, sizing.arena_floats_per_frame);
                if (plan) {  // the stages (= kernel launches per chunk) this segment of this group lowers to
                    std::vector<int> kinds;
                    for (auto& kv : sizing.builds) kinds.push_back(kv.second.kind);
                    so[k].stage_lists.push_back(std::move(kinds));
                    if (plan_digest_wanted()) so[k].digest = digest_builds(sizing.builds, so[k].digest);
                    if (check_split && one_segment) serial_digest = digest_builds(sizing.builds, 1469598103934665603ull);
                }
            }
            b->groups[k].src_floats = sizing.src_cursor;
            so[k].has_feedback = sizing.has_feedback;
            so[k].delay_ch_seen = std::move(sizing.delay_ch_seen);
            if (plan && check_split && one_segment && n_in_group >= 2) {  // the split sizing of the same group must say the same
                RangeOut out;
                // (WAE_PLAN_PARALLEL=2: the runs on real worker threads, as the one-shot render sizes them — for the sanitizers)
                static const bool threaded = [] { const char* e = getenv("WAE_PLAN_PARALLEL"); return e && atoi(e) >= 2; }();
                std::unique_ptr<WorkerPool> tmp_pool(threaded ? new WorkerPool(3, 0) : nullptr);
                const bool ok = size_group_split(k, tmp_pool.get(), (int)std::min<uint32_t>(n_in_group, 3u), out);
                bool same = ok && out.fpf == so[k].fpf && out.src_floats == b->groups[k].src_floats && out.has_feedback == so[k].has_feedback &&
                            out.delay_ch_seen == so[k].delay_ch_seen && out.copies.size() == b->groups[k].src_copies.size() &&
                            out.graph_base == b->groups[k].graph_src_base;
                if (same) {  // with the runs started at their place in the slab the merged stage builds ARE the serial ones
                    RangeOut abs_out;
                    same = size_group_split(k, nullptr, (int)std::min<uint32_t>(n_in_group, 3u), abs_out, &b->groups[k].graph_src_base) &&
                           abs_out.graph_base == b->groups[k].graph_src_base && digest_builds(abs_out.builds, 1469598103934665603ull) == serial_digest;
                }
                for (size_t c = 0; same && c < out.copies.size(); c++) {
                    const auto& x = out.copies[c];
                    const auto& y = b->groups[k].src_copies[c];
                    same = x.buf == y.buf && x.offset == y.offset && x.floats == y.floats;
                }
                if (!same) {
                    so[k].code = WAE_INVALID_STATE;
                    so[k].error = "internal: the split sizing pass disagrees with the serial one";
                }
            }
        };
        if (pool) pool->parallel_for(n_groups, size_group);
        else
            for (int k = 0; k < n_groups; k++) size_group(k);
        for (int k = 0; k < n_groups; k++) {
            if (so[k].code != WAE_OK) {
                int code = so[k].code;
                std::string msg = so[k].error;
                wae_batch_destroy(b);
                return fail(code, msg);
            }
            fpf = std::max(fpf, so[k].fpf);
            has_feedback = has_feedback || so[k].has_feedback;
            for (auto& l : so[k].stage_lists) plan_stage_lists.push_back(std::move(l));
            if (plan && plan_digest_wanted()) std::fprintf(stderr, "[wae plan digest] group %d fpf %llu src %zu: %016llx\n", k, (unsigned long long)so[k].fpf, b->groups[k].src_floats, (unsigned long long)so[k].digest);
            for (auto& kv : so[k].delay_ch_seen) {
                auto it = ps.delay_ch_hint.find(kv.first);
                int cur = it == ps.delay_ch_hint.end() ? 1 : it->second;
                if (cur != kv.second) {
                    ps.delay_ch_hint[kv.first] = kv.second;
                    hints_changed = true;
                }
            }
        }
        if (!hints_changed) {
            converged = true;
            break;
        }
    }
    if (!converged) {
        wae_batch_destroy(b);
        return fail(WAE_UNSUPPORTED, "the channel layout of DelayNodes inside feedback cycles did not settle (graph not lowered to the GPU)");
    }
    ps.t1 = std::chrono::steady_clock::now();  // sizing pass done
    b->arena_bytes = 0;
    b->asset_bytes = 0;
    int64_t chunk = eng->chunk_frames;
    if (chunk == 0) {
        if (fpf == 0) {
            chunk = b->lq;
        } else {
            chunk = (int64_t)(1024.0 * 1024 * 1024 / (4.0 * (double)fpf));  // arena budget 1 GiB
            chunk = std::max<int64_t>(8192, std::min<int64_t>(chunk, 1 << 20));  // >= 8192: keeps per-chunk launches and serial tails amortised
            chunk = chunk / 2048 * 2048;
        }
        if (has_conv || has_hrtf) chunk = std::max<int64_t>(chunk, 8 * WAE_CONV_BLOCK);  // k_conv_mac tiles 8 output blocks
    }
    // feedback through a DelayNode: the stages up to the cycle are replayed one render quantum at a time INSIDE each chunk
    // (run_group), so the chunk size does not depend on it
    (void)has_feedback;
    if (has_conv || has_hrtf) chunk = (chunk + WAE_CONV_BLOCK - 1) / WAE_CONV_BLOCK * WAE_CONV_BLOCK;
    if (chunk > b->lq) chunk = (has_conv || has_hrtf) ? (b->lq + WAE_CONV_BLOCK - 1) / WAE_CONV_BLOCK * WAE_CONV_BLOCK : b->lq;
    b->chunk = chunk;
    if (plan) {
        std::memset(plan, 0, sizeof *plan);
        plan->groups = (uint32_t)n_groups;
        plan->chunk_frames = (uint64_t)chunk;
        plan->chunks = (uint64_t)((b->lq + chunk - 1) / chunk);
        plan->arena_floats_per_frame = fpf;
        plan->has_feedback = has_feedback ? 1u : 0u;
        uint32_t per_kind[S_KINDS] = {0};
        for (auto& grp : b->groups) {
            plan->source_floats += grp.src_floats;
            plan->segments += (uint32_t)grp.seg_bounds.size() - 1;
        }
        for (auto& kinds : plan_stage_lists) {
            plan->stages += (uint32_t)kinds.size();
            for (int k : kinds) per_kind[k]++;
        }
        std::string txt;  // "k_chain x 1, k_mix x 1": stages by kernel, summed over groups and segments
        for (int k = 0; k < S_KINDS; k++)
            if (per_kind[k]) txt += (txt.empty() ? "" : ", ") + std::string(kStageNames[k]) + " x " + std::to_string(per_kind[k]);
        std::snprintf(plan->stage_kinds, sizeof plan->stage_kinds, "%s", txt.c_str());
        delete b;  // nothing was allocated on the device
        *out_b = nullptr;
        return WAE_OK;
    }
    for (auto& grp : b->groups) {
        if (cudaEventCreateWithFlags(&grp.ev_h2d, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&grp.ev_done, cudaEventDisableTiming) != cudaSuccess) {
            const std::string msg = std::string("prepare: cudaEventCreate: ") + cudaGetErrorString(cudaGetLastError());
            wae_batch_destroy(b);
            return fail(WAE_CUDA_ERROR, msg);
        }
        if (grp.src_floats) {
            // (not zeroed: every float of the slab is covered by a source copy, channel paddings included)
            grp.d_src = b->dalloc<float>(grp.src_floats, false);
            if (!grp.d_src) {
                wae_batch_destroy(b);
                return fail(WAE_OUT_OF_MEMORY, "out of device memory (source PCM slab)");
            }
        }
    }
    *out_b = b;
    return WAE_OK;
}

// phase B: plans group k (all its segments) and uploads its tables.  Thread-safe against other groups of the same batch.
static void prep_plan_group(wae_batch* b, wae_graph* const* graphs, int k, PrepState& ps, GroupPlan& gp) {
    wae_engine* eng = b->engine;
    wae_batch::Group& grp = b->groups[k];
    Planner pl{b, eng};
    pl.d_src = grp.d_src;
    pl.group_graphs = (int)(grp.g1 - grp.g0);
    pl.src_copies = nullptr;  // recorded by the sizing pass
    pl.delay_ch_hint = &ps.delay_ch_hint;
    pl.ir_cache = &ps.ir_cache;
    auto oom = [&](const char* what) {
        gp.code = WAE_OUT_OF_MEMORY;
        gp.error = std::string("out of device memory (") + what + ")";
    };
    // The only group of a batch of few, large graphs, no suspend points (north_star: 8 graphs of 3000 nodes): planned by several workers,
    // one Planner per contiguous run of graphs starting at the slab cursor the sizing pass recorded for its first graph; the runs' stage
    // builds are merged in graph order (merge_builds) — the tables a single planner would have built (wae_batch_plan checks that on the
    // CPU under WAE_PLAN_PARALLEL=1).  This function may itself run on a worker (one-shot render), which then waits for the others: only
    // with one group in the batch, and never with fewer than two workers left.
    int split_parts = 0;
    {
        static const bool split_on = [] { const char* e = getenv("WAE_PLAN_SPLIT"); return !e || atoi(e) != 0; }();
        const uint32_t n_in_group = grp.g1 - grp.g0;
        size_t nn = 0;
        for (uint32_t i = grp.g0; i < grp.g1; i++) nn += graphs[i]->nodes.size();
        if (split_on && b->groups.size() == 1 && grp.seg_bounds.size() == 2 && n_in_group >= 2 && nn >= 4096 && grp.graph_src_base.size() == n_in_group) {
            const int free_workers = eng->workers()->size() - 1;
            if (free_workers >= 2) split_parts = (int)std::min<uint32_t>(n_in_group, (uint32_t)free_workers);
        }
    }
    for (int sg = 0; sg + 1 < (int)grp.seg_bounds.size(); sg++) {
        pl.begin_segment(grp.seg_bounds[sg], grp.seg_bounds[sg + 1]);
        const uint64_t alg_before = pl.algorithmic_bytes;
        if (split_parts >= 2) {
            struct Run {
                Builds builds;
                uint64_t algorithmic_bytes = 0;
                int code = WAE_OK;
                std::string error;
            };
            std::vector<Run> runs((size_t)split_parts);
            const uint32_t n = grp.g1 - grp.g0;
            eng->workers()->parallel_for(split_parts, [&](int t) {
                const uint32_t i0 = (uint32_t)((uint64_t)n * t / split_parts), i1 = (uint32_t)((uint64_t)n * (t + 1) / split_parts);
                Planner rp{b, eng};
                rp.d_src = grp.d_src;
                rp.group_graphs = (int)n;
                rp.src_copies = nullptr;
                rp.delay_ch_hint = &ps.delay_ch_hint;
                rp.ir_cache = &ps.ir_cache;
                rp.src_cursor = grp.graph_src_base[i0];
                rp.begin_segment(grp.seg_bounds[sg], grp.seg_bounds[sg + 1]);
                for (uint32_t i = grp.g0 + i0; i < grp.g0 + i1; i++) {
                    EpochView view(graphs[i], grp.seg_bounds[sg]);
                    if (!rp.plan_graph(graphs[i], i)) {
                        runs[t].code = rp.error_code;
                        runs[t].error = rp.error;
                        return;
                    }
                }
                runs[t].builds = std::move(rp.builds);
                runs[t].algorithmic_bytes = rp.algorithmic_bytes;
            });
            for (auto& r : runs) {
                if (r.code != WAE_OK) {
                    gp.code = r.code;
                    gp.error = r.error;
                    return;
                }
                merge_builds(pl.builds, r.builds);
                pl.algorithmic_bytes += r.algorithmic_bytes;
            }
        } else
        for (uint32_t i = grp.g0; i < grp.g1; i++) {
            EpochView view(graphs[i], grp.seg_bounds[sg]);
            if (!pl.plan_graph(graphs[i], i)) {
                gp.code = pl.error_code;
                gp.error = pl.error;
                return;
            }
        }
        // the per-node byte counts assume the whole render: scale to this segment's share of it
        gp.algorithmic_bytes += (uint64_t)((double)(pl.algorithmic_bytes - alg_before) * (double)(pl.seg_end - pl.seg_start) / (double)b->lq);
        // materialise the segment's stages in (class, level, kind) order
        const size_t seg_stage0 = gp.stages.size();
        void* last_conv_inputs = nullptr;
        for (auto& kv : pl.builds) {
            StageBuild& s = kv.second;
            Stage st;
            st.seg = sg;
            st.cls = s.cls;
            st.kind = s.kind;
            st.variant = s.variant;
            st.group = k;
            st.max_ch = s.max_ch;
            switch (s.kind) {
                case S_MIX: {
                    for (auto& m : s.mix) {  // classify: vector fast path of k_mix
                        bool simple = true, all_mono = m.n_edges > 0;
                        for (int e = 0; e < m.n_edges; e++) {
                            const MixEdge& ed = s.mix_edges[m.edge_offset + e];
                            bool same = ed.src_ch == m.out_ch;
                            bool dup = ed.src_ch == 1 && m.out_ch == 2 && m.interp == WAE_INTERPRETATION_SPEAKERS;
                            if (!(same || dup)) simple = false;
                            if (ed.src_ch != 1) all_mono = false;
                            // sources must allow 16-byte loads: arena buffers do; asset / output aliases may not
                            if (ed.src.absolute || (ed.src.stride & 3) != 0 || (reinterpret_cast<uintptr_t>(ed.src.p) & 15) != 0) simple = false;
                        }
                        m.simple = simple ? 1 : 0;
                        m.all_mono = (simple && all_mono && (m.out_ch == 1 || (m.out_ch == 2 && m.interp == WAE_INTERPRETATION_SPEAKERS))) ? 1 : 0;
                    }
                    st.n = (int)s.mix.size(); st.d_a = up(b, s.mix); st.d_b = up(b, s.mix_edges);
                    for (auto& m : s.mix) st.n_b = std::max(st.n_b, (int)m.n_edges);  // widest port: picks the mixer kernel
                    break;
                }
                case S_MIX_DYN: {
                    for (auto& m : s.mix_dyn) {  // classify: the four-frames-per-thread path of k_mix_dyn
                        bool ok = m.out_ch <= 2;
                        for (int e = 0; e < m.n_edges && ok; e++) {
                            const MixEdge& ed = s.mix_edges[m.edge_offset + e];
                            if (ed.src_ch > 2 || ed.src.absolute || (ed.src.stride & 3) != 0 || (reinterpret_cast<uintptr_t>(ed.src.p) & 15) != 0) ok = false;
                        }
                        m.stereo4 = ok ? 1 : 0;
                    }
                    st.n = (int)s.mix_dyn.size(); st.d_a = up(b, s.mix_dyn); st.d_b = up(b, s.mix_edges);
                    break;
                }
                case S_META: st.n = (int)s.meta.size(); st.d_a = up(b, s.meta); break;
                case S_OSC: st.n = (int)s.osc.size(); st.d_a = up(b, s.osc); break;
                case S_CONST: st.n = (int)s.cst.size(); st.d_a = up(b, s.cst); break;
                case S_ABSN: st.n = (int)s.absn.size(); st.d_a = up(b, s.absn); break;
                case S_BIQUAD: st.n = (int)s.biquad.size(); st.d_a = up(b, s.biquad); break;
                case S_CHAIN: {
                    st.n = (int)s.chain.size(); st.d_a = up(b, s.chain); st.d_b = up(b, s.scan_coef);
                    const int nb = (s.variant % 6) / 2;
                    int slabs = 1, tps = 1;
                    int pre_log2 = -1;
                    if (nb > 0) chain_plan_slabs(st.n, st.max_ch, (int)std::min<int64_t>(b->chunk, pl.seg_end - pl.seg_start), nb, &slabs, &tps, &pre_log2);
                    if (slabs > 1) {  // time slabs of filtered chains hand their state over through device memory
                        const size_t slots = (size_t)st.n * st.max_ch * slabs;
                        st.chain.slab_stride = slabs;
                        st.chain.ticket = b->dalloc<unsigned>(1, true);
                        st.chain.flags = b->dalloc<unsigned>(slots, true);
                        st.chain.handoff = b->dalloc<double>(slots * CHAIN_MAX_BIQUADS * 4);
                        if (!st.chain.ticket || !st.chain.flags || !st.chain.handoff) return oom("chain hand-off");
                    }
                    break;
                }
                case S_VSUM: {
                    st.n = (int)s.vgroups.size(); st.d_a = up(b, s.chain); st.d_b = up(b, s.scan_coef); st.d_c = up(b, s.vgroups);
                    const int64_t nf_max = std::min<int64_t>(b->chunk, pl.seg_end - pl.seg_start);
                    const int tiles = (int)((nf_max + 2047) / 2048);
                    st.chain.slab_stride = tiles;  // progress counters per group
                    st.chain.ticket = b->dalloc<unsigned>(1, true);
                    st.chain.flags = b->dalloc<unsigned>((size_t)st.n * tiles, true);
                    st.chain.handoff = b->dalloc<double>(s.chain.size() * 2 * 4);
                    if (!st.chain.ticket || !st.chain.flags || !st.chain.handoff) return oom("voice-sum hand-off");
                    break;
                }
                case S_PARAM: st.n = (int)s.param.size(); st.d_a = up(b, s.param); break;
                case S_OSC_AR: st.n = (int)s.osc_ar.size(); st.d_a = up(b, s.osc_ar); break;
                case S_BIQUAD_AR: st.n = (int)s.biquad_ar.size(); st.d_a = up(b, s.biquad_ar); break;
                case S_ABSN_SLOW: st.n = (int)s.absn_slow.size(); st.d_a = up(b, s.absn_slow); break;
                case S_IIR: st.n = (int)s.iir.size(); st.d_a = up(b, s.iir); break;
                case S_GAIN: st.n = (int)s.gain.size(); st.d_a = up(b, s.gain); break;
                case S_SHAPER: st.n = (int)s.shaper.size(); st.d_a = up(b, s.shaper); break;
                case S_SPAN: st.n = (int)s.span.size(); st.d_a = up(b, s.span); st.d_b = up(b, s.span_gains); break;
                case S_PAN: st.n = (int)s.pan.size(); st.d_a = up(b, s.pan); break;
                case S_HRTF: st.n = (int)s.hrtf.size(); st.d_a = up(b, s.hrtf); st.max_ch = s.hrtf.empty() ? 0 : s.hrtf[0].L;
                    st.n_b = (int)s.hrtf_sel.size(); st.d_b = up(b, s.hrtf_sel); break;
                case S_PAN_DYN: st.n = (int)s.pan_dyn.size(); st.d_a = up(b, s.pan_dyn); break;
                case S_ABSN_SERIAL: st.n = (int)s.absn_serial.size(); st.d_a = up(b, s.absn_serial); break;
                case S_SHAPER_OS: st.n = (int)s.shaper_os.size(); st.d_a = up(b, s.shaper_os); break;
                case S_ROUTE: st.n = (int)s.route.size(); st.d_a = up(b, s.route); break;
                case S_DELAY_MONO:
                case S_DELAY:
                case S_DELAY_WRITE: st.n = (int)s.delay.size(); st.d_a = up(b, s.delay); break;
                case S_COMP: st.n = (int)s.comp.size(); st.d_a = up(b, s.comp); break;
                case S_ANALYSER: st.n = (int)s.analyser.size(); st.d_a = up(b, s.analyser); break;
                case S_CONV_FFT: st.n = (int)s.conv_in.size(); st.d_a = up(b, s.conv_in); last_conv_inputs = st.d_a; break;
                case S_CONV_MAC:
                case S_CONV_MAC_ACC:
                    st.n = (int)s.conv_path.size();
                    st.d_a = up(b, s.conv_path);
                    st.d_b = last_conv_inputs;  // conv-input table of the same level (kinds are ordered FFT < MAC < MAC_ACC)
                    break;
            }
            if (st.n > 0) {
                if (!st.d_a) return oom("stage tables");
                gp.stages.push_back(st);
            }
        }
        gp.seg_ranges.push_back({seg_stage0, gp.stages.size()});
    }  // segments
    b->flush_uploads();  // the group's tables, before anything of it is launched
}

static void prep_append_group(wae_batch* b, int k, PrepState& ps, GroupPlan& gp) {
    wae_batch::Group& grp = b->groups[k];
    grp.stage0 = b->stages.size();
    for (auto& r : gp.seg_ranges) grp.seg_stages.push_back({grp.stage0 + r.first, grp.stage0 + r.second});
    for (auto& st : gp.stages) b->stages.push_back(st);
    grp.stage1 = b->stages.size();
    ps.algorithmic_bytes += gp.algorithmic_bytes;
}

// H2D of a group's source PCM, one copy per AudioBufferSourceNode, straight from the buffers the graphs own
static wae_status enqueue_source_copies(wae_batch* b, wae_batch::Group& grp, cudaStream_t s) {
    for (auto& sc : grp.src_copies)
        CUDA_TRY(cudaMemcpyAsync(grp.d_src + sc.offset, sc.buf->base, sc.floats * sizeof(float), cudaMemcpyHostToDevice, s));
    return WAE_OK;
}

static wae_status prep_finish(wae_batch* b, PrepState& ps) {
    CUDA_TRY(cudaEventCreate(&b->ev0));
    CUDA_TRY(cudaEventCreate(&b->ev1));
    int64_t n_chunks = 0;
    for (auto& grp : b->groups)  // the largest number of chunks any group renders
    {
        int64_t c = 0;
        for (size_t sg = 0; sg + 1 < grp.seg_bounds.size(); sg++) c += (grp.seg_bounds[sg + 1] - grp.seg_bounds[sg] + b->chunk - 1) / b->chunk;
        n_chunks = std::max(n_chunks, c);
    }
    uint64_t launches = 0;
    for (auto& st : b->stages) {
        const uint64_t k = (st.kind == S_CONV_FFT || st.kind == S_CONV_MAC || st.kind == S_CONV_MAC_ACC || st.kind == S_SHAPER_OS) ? 2
                           : st.kind == S_HRTF ? (st.n_b > 0 ? 3 : 2) : 1;
        // per-quantum stages (class 1) launch once per render quantum of their segment, the others once per chunk of it
        const std::vector<int64_t>& sb = b->groups[st.group].seg_bounds;
        const int64_t seg_len = sb[st.seg + 1] - sb[st.seg];
        launches += st.cls == 1 ? k * (uint64_t)(seg_len / 128) : k * (uint64_t)((seg_len + b->chunk - 1) / b->chunk);
    }
    std::memset(&b->stats, 0, sizeof(b->stats));
    b->stats.kernel_launches_per_run = launches;
    b->stats.stages = b->stages.size();
    b->stats.chunks = (uint64_t)n_chunks;
    b->stats.arena_bytes = b->arena_bytes;
    b->stats.asset_bytes = b->asset_bytes;
    b->stats.algorithmic_bytes = ps.algorithmic_bytes;
    b->stats.graph_quanta = (uint64_t)b->n_graphs * (uint64_t)(b->lq / 128);
    return WAE_OK;
}

// `plan` != nullptr: planning only — grouping, the sizing pass of the planner (which touches no device memory) and the chunk choice,
// reported through *plan; nothing is allocated and no CUDA call is made (wae_batch_plan: runs without a GPU).
static wae_status prepare_impl(wae_engine* eng, wae_graph* const* graphs, uint32_t n_graphs, wae_batch** out, wae_plan_info* plan) {
    if (!out && !plan) return fail(WAE_INVALID_ARGUMENT, "null out pointer");
    PrepState ps;
    wae_batch* b = nullptr;
    wae_status st = prep_begin(eng, graphs, n_graphs, plan, &b, ps, true);
    if (st != WAE_OK || plan) return st;
    const int n_groups = (int)b->groups.size();
    std::vector<GroupPlan> gps(n_groups);
    WorkerPool* pool = (n_groups > 1 && n_graphs >= 64) ? eng->workers() : nullptr;
    if (pool) pool->parallel_for(n_groups, [&](int k) { prep_plan_group(b, graphs, k, ps, gps[k]); });
    else
        for (int k = 0; k < n_groups; k++) prep_plan_group(b, graphs, k, ps, gps[k]);
    for (int k = 0; k < n_groups; k++) {
        if (gps[k].code != WAE_OK) {
            int code = gps[k].code;
            std::string msg = gps[k].error;
            wae_batch_destroy(b);
            return fail(code, msg);
        }
        prep_append_group(b, k, ps, gps[k]);
        // first upload of the group's source PCM, straight from the graphs' buffers (page-locked when the graphs have an engine)
        wae_status cs = enqueue_source_copies(b, b->groups[k], eng->stream);
        if (cs != WAE_OK) {
            wae_batch_destroy(b);
            return cs;
        }
    }
    const auto t_prep2 = std::chrono::steady_clock::now();  // planned, allocated, uploads enqueued
    st = prep_finish(b, ps);
    if (st == WAE_OK && cudaStreamSynchronize(eng->stream) != cudaSuccess) st = fail(WAE_CUDA_ERROR, "prepare: stream synchronisation failed");
    if (st != WAE_OK) {
        wae_batch_destroy(b);
        return st;
    }
    if (getenv("WAE_PREPARE_PROFILE")) {
        const auto t_prep3 = std::chrono::steady_clock::now();
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point c) { return std::chrono::duration<double, std::milli>(c - a).count(); };
        std::fprintf(stderr, "[wae prepare] sizing %.1f ms, plan+alloc+upload %.1f ms, sync %.1f ms, fresh device blocks %llu, stages %zu\n", ms(ps.t0, ps.t1),
                     ms(ps.t1, t_prep2), ms(t_prep2, t_prep3), (unsigned long long)b->n_cuda_malloc, b->stages.size());
    }
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) {
        wae_batch_destroy(b);
        return fail(WAE_CUDA_ERROR, std::string("prepare: ") + cudaGetErrorString(le));
    }
    *out = b;
    return WAE_OK;
}

WAE_API wae_status wae_batch_prepare(wae_engine* eng, wae_graph* const* graphs, uint32_t n_graphs, wae_batch** out) {
    if (!out) return fail(WAE_INVALID_ARGUMENT, "null out pointer");
    return prepare_impl(eng, graphs, n_graphs, out, nullptr);
}

// What wae_batch_prepare would lower the graphs to, without a device: the planner's sizing pass under the default engine options.
WAE_API wae_status wae_batch_plan(wae_graph* const* graphs, uint32_t n_graphs, wae_plan_info* info) {
    if (!info) return fail(WAE_INVALID_ARGUMENT, "null info pointer");
    wae_engine host_only;  // default options; no stream, no sphere: HRTF panners answer WAE_UNSUPPORTED ("needs an HRIR sphere")
    return prepare_impl(&host_only, graphs, n_graphs, nullptr, info);
}

static void launch_stage(wae_batch* b, Stage& st, ChunkInfo ci) {
    cudaStream_t s = b->engine->stream;
    switch (st.kind) {
        case S_MIX: launch_mix((MixInst*)st.d_a, (MixEdge*)st.d_b, st.n, ci, s, st.n_b); break;
        case S_MIX_DYN: launch_mix_dyn((MixDynInst*)st.d_a, (MixEdge*)st.d_b, st.n, ci, s); break;
        case S_META: launch_meta((MetaInst*)st.d_a, st.n, ci, s); break;
        case S_DELAY_MONO: launch_delay_mono((DelayInst*)st.d_a, st.n, ci, s); break;
        case S_OSC: launch_oscillator((OscInst*)st.d_a, st.n, ci, s); break;
        case S_CONST: launch_constant((ConstInst*)st.d_a, st.n, ci, s); break;
        case S_ABSN: launch_buffer_source((AbsnInst*)st.d_a, st.n, ci, s); break;
        case S_BIQUAD:
            launch_biquad_serial((BiquadInst*)st.d_a, st.n, st.max_ch, ci, s);
            break;
        case S_PARAM: launch_param((ParamInst*)st.d_a, st.n, ci, s, b->engine->param_parallel); break;
        case S_OSC_AR: launch_osc_arate((OscArInst*)st.d_a, st.n, ci, s); break;
        case S_ABSN_SLOW: launch_buffer_source_slow((AbsnSlowInst*)st.d_a, st.n, ci, s); break;
        case S_BIQUAD_AR: launch_biquad_arate((BiquadArInst*)st.d_a, st.n, st.max_ch, ci, s); break;
        case S_CHAIN:
            st.chain.epoch++;  // hand-off flags of this launch carry its number (never reset, never reused)
            launch_chain(st.variant, (ChainInst*)st.d_a, (ScanCoef*)st.d_b, st.n, st.max_ch, ci, s, st.chain);
            break;
        case S_VSUM:
            st.chain.epoch++;  // (progress counters carry the launch number)
            launch_voice_sum(st.variant, (ChainInst*)st.d_a, (ScanCoef*)st.d_b, (VoiceGroup*)st.d_c, st.n, ci, s, st.chain);
            break;
        case S_IIR: launch_iir((IirInst*)st.d_a, st.n, st.max_ch, ci, s); break;
        case S_GAIN: launch_gain((GainInst*)st.d_a, st.n, ci, s); break;
        case S_SHAPER: launch_shaper((ShaperInst*)st.d_a, st.n, ci, s); break;
        case S_SPAN: launch_stereo_panner((SPanInst*)st.d_a, (float2*)st.d_b, st.n, ci, s); break;
        case S_PAN: launch_panner_eq((PanInst*)st.d_a, st.n, ci, s); break;
        case S_HRTF: launch_hrtf((HrtfInst*)st.d_a, st.n, (HrtfSelInst*)st.d_b, st.n_b, st.max_ch, ci, s); break;
        case S_PAN_DYN: launch_panner_dyn((PanDynInst*)st.d_a, st.n, ci, s); break;
        case S_ABSN_SERIAL: launch_buffer_source_serial((AbsnSerialInst*)st.d_a, st.n, ci, s); break;
        case S_SHAPER_OS: launch_shaper_os((ShaperOsInst*)st.d_a, st.n, st.max_ch, ci, s); break;
        case S_ROUTE: launch_route((RouteInst*)st.d_a, st.n, ci, s); break;
        case S_DELAY: launch_delay_read((DelayInst*)st.d_a, st.n, ci, s); break;
        case S_DELAY_WRITE: launch_ring_write((DelayInst*)st.d_a, st.n, ci, s); break;
        case S_COMP: launch_compressor((CompInst*)st.d_a, st.n, ci, s); break;
        case S_ANALYSER: launch_analyser((AnalyserInst*)st.d_a, st.n, ci, s); break;
        case S_CONV_FFT: launch_conv_fft_in((ConvInput*)st.d_a, st.n, ci, s); break;
        case S_CONV_MAC:
        case S_CONV_MAC_ACC: launch_conv_mac_ifft((ConvPath*)st.d_a, (ConvInput*)st.d_b, st.n, ci, s); break;
    }
}

// Groups whose source PCM is not all page-locked (graphs built without an engine, buffers below the pinning threshold) get a pinned
// mirror of their slab, built once, when the PCM is uploaded a second time; page-locked buffers are copied from where they are.
static bool group_sources_pinned(const wae_batch::Group& g) {
    for (auto& sc : g.src_copies)
        if (!sc.buf->pinned) return false;
    return true;
}
static wae_status ensure_host_mirror(wae_batch* b) {
    for (auto& g : b->groups) {
        if (!g.src_floats || g.h_src || group_sources_pinned(g)) continue;
        void* hp = nullptr;
        if (cudaHostAlloc(&hp, g.src_floats * sizeof(float), cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return fail(WAE_OUT_OF_MEMORY, "out of memory (pinned mirror of the source PCM)");
        }
        b->pinned.push_back(hp);
        g.h_src = (float*)hp;
        for (auto& sc : g.src_copies) std::memcpy(g.h_src + sc.offset, sc.buf->base, sc.floats * sizeof(float));
    }
    return WAE_OK;
}
static wae_status resend_group_sources(wae_batch* b, wae_batch::Group& g, cudaStream_t s) {
    if (!g.src_floats) return WAE_OK;
    if (g.h_src) {
        CUDA_TRY(cudaMemcpyAsync(g.d_src, g.h_src, g.src_floats * sizeof(float), cudaMemcpyHostToDevice, s));
        return WAE_OK;
    }
    return enqueue_source_copies(b, g, s);
}

// re-upload the source PCM of every AudioBufferSourceNode (page-locked buffers / pinned mirror)
WAE_API wae_status wae_batch_upload(wae_batch* b) {
    CUDA_TRY(cudaSetDevice(b->engine->device));
    wae_status ms = ensure_host_mirror(b);
    if (ms != WAE_OK) return ms;
    for (auto& g : b->groups) {
        ms = resend_group_sources(b, g, b->engine->stream);
        if (ms != WAE_OK) return ms;
    }
    return WAE_OK;
}

WAE_API wae_status wae_batch_set_timing(wae_batch* b, uint32_t per_stage) {
    b->time_stages = per_stage != 0;
    return WAE_OK;
}

// renders one group (all its chunks, all its stages) on the engine stream
static wae_status run_group(wae_batch* b, const wae_batch::Group& g) {
    cudaStream_t s = b->engine->stream;
    // per-stage device time: one event between consecutive launches, recorded on the launching stream and read back in
    // wae_batch_sync (no host synchronisation inside the run)
    size_t e_prev = (size_t)-1;
    auto next_event = [&]() -> size_t {
        if (b->timed_events_used == b->stage_events.size()) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return (size_t)-1;
            b->stage_events.push_back(e);
        }
        return b->timed_events_used++;
    };
    auto run = [&](size_t i, const ChunkInfo& ci) -> bool {
        launch_stage(b, b->stages[i], ci);
        if (b->time_stages) {
            size_t e = next_event();
            if (e == (size_t)-1) return false;
            if (cudaEventRecord(b->stage_events[e], s) != cudaSuccess) return false;
            b->timed.push_back({i, e_prev, e});
            e_prev = e;
        }
        return true;
    };
    for (size_t sg = 0; sg + 1 < g.seg_bounds.size(); sg++) {  // render segments between suspend points (usually one)
        const size_t s0 = g.seg_stages[sg].first, s1 = g.seg_stages[sg].second;
        // stages are sorted by class: [whole-chunk stages of feedback-free graphs | per-quantum stages | whole-chunk stages after cycles]
        size_t c1 = s0, c2 = s0;
        while (c1 < s1 && b->stages[c1].cls == 0) c1++;
        c2 = c1;
        while (c2 < s1 && b->stages[c2].cls == 1) c2++;
        const int64_t seg_end = g.seg_bounds[sg + 1];
        for (int64_t f0 = g.seg_bounds[sg]; f0 < seg_end; f0 += b->chunk) {
            const ChunkInfo ci{f0, (int32_t)std::min<int64_t>(b->chunk, seg_end - f0), 0};
            if (b->time_stages) {
                e_prev = next_event();
                if (e_prev == (size_t)-1) return fail(WAE_CUDA_ERROR, "cudaEventCreate failed");
                CUDA_TRY(cudaEventRecord(b->stage_events[e_prev], s));
            }
            for (size_t i = s0; i < c1; i++)
                if (!run(i, ci)) return fail(WAE_CUDA_ERROR, "cudaEventRecord failed");
            if (c2 > c1)
                for (int32_t sub = 0; sub < ci.nf; sub += 128) {  // the reference's render loop, for the cyclic part only
                    const ChunkInfo cq{f0 + sub, 128, sub};
                    for (size_t i = c1; i < c2; i++)
                        if (!run(i, cq)) return fail(WAE_CUDA_ERROR, "cudaEventRecord failed");
                }
            for (size_t i = c2; i < s1; i++)
                if (!run(i, ci)) return fail(WAE_CUDA_ERROR, "cudaEventRecord failed");
        }
    }
    return WAE_OK;
}

static wae_status begin_run(wae_batch* b) {
    CUDA_TRY(cudaSetDevice(b->engine->device));
    cudaStream_t s = b->engine->stream;
    for (auto& z : b->zero_on_run) CUDA_TRY(cudaMemsetAsync(z.first, 0, z.second, s));
    b->timed.clear();
    b->timed_events_used = 0;
    for (auto& a : b->analysers) a.computed = false;
    CUDA_TRY(cudaEventRecord(b->ev0, s));
    return WAE_OK;
}

WAE_API wae_status wae_batch_run(wae_batch* b) {
    wae_status st = begin_run(b);
    if (st != WAE_OK) return st;
    for (auto& g : b->groups) {
        st = run_group(b, g);
        if (st != WAE_OK) return st;
    }
    CUDA_TRY(cudaEventRecord(b->ev1, b->engine->stream));
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) return fail(WAE_CUDA_ERROR, std::string("run: ") + cudaGetErrorString(le));
    return WAE_OK;
}

WAE_API wae_status wae_batch_group_count(wae_batch* b, uint32_t* n_groups) {
    if (!b || !n_groups) return fail(WAE_INVALID_ARGUMENT, "null argument");
    *n_groups = (uint32_t)b->groups.size();
    return WAE_OK;
}
WAE_API wae_status wae_batch_group_range(wae_batch* b, uint32_t group, uint32_t* first_graph, uint32_t* last_graph) {
    if (!b || !first_graph || !last_graph || group >= b->groups.size()) return fail(WAE_INVALID_ARGUMENT, "null argument / group out of range");
    *first_graph = b->groups[group].g0;
    *last_graph = b->groups[group].g1;
    return WAE_OK;
}
WAE_API wae_status wae_batch_run_group(wae_batch* b, uint32_t group) {
    if (!b || group >= b->groups.size()) return fail(WAE_INVALID_ARGUMENT, "null batch / group out of range");
    wae_status st = WAE_OK;
    if (group == 0) st = begin_run(b);
    else CUDA_TRY(cudaSetDevice(b->engine->device));
    if (st != WAE_OK) return st;
    st = run_group(b, b->groups[group]);
    if (st != WAE_OK) return st;
    if (group + 1 == b->groups.size()) CUDA_TRY(cudaEventRecord(b->ev1, b->engine->stream));
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) return fail(WAE_CUDA_ERROR, std::string("run_group: ") + cudaGetErrorString(le));
    return WAE_OK;
}

// End-to-end render with HOST buffers: for every group, H2D of its source PCM (pinned mirror), render, D2H of its
// rendered PCM into `host_out` ([n_graphs][channels][length] f32; pinned memory gives full PCIe speed) — on three
// streams, so the copies of neighbouring groups overlap the render.  Synchronous: returns when host_out is complete.
static wae_status run_pipelined(wae_batch* b, float* host_out, bool resend_sources) {
    wae_status st = resend_sources ? ensure_host_mirror(b) : WAE_OK;
    if (st != WAE_OK) return st;
    st = begin_run(b);
    if (st != WAE_OK) return st;
    cudaStream_t s = b->engine->stream;
    const size_t per_graph = (size_t)b->channels * b->length;
    for (auto& g : b->groups) {
        if (g.src_floats && resend_sources) {
            st = resend_group_sources(b, g, b->s_h2d);
            if (st != WAE_OK) return st;
            CUDA_TRY(cudaEventRecord(g.ev_h2d, b->s_h2d));
            CUDA_TRY(cudaStreamWaitEvent(s, g.ev_h2d, 0));
        }
        st = run_group(b, g);
        if (st != WAE_OK) return st;
        CUDA_TRY(cudaEventRecord(g.ev_done, s));
        CUDA_TRY(cudaStreamWaitEvent(b->s_d2h, g.ev_done, 0));
        CUDA_TRY(cudaMemcpyAsync(host_out + (size_t)g.g0 * per_graph, b->d_out + (size_t)g.g0 * per_graph,
                                 (size_t)(g.g1 - g.g0) * per_graph * sizeof(float), cudaMemcpyDeviceToHost, b->s_d2h));
    }
    CUDA_TRY(cudaEventRecord(b->ev1, s));
    CUDA_TRY(cudaStreamSynchronize(b->s_d2h));
    CUDA_TRY(cudaStreamSynchronize(s));
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) return fail(WAE_CUDA_ERROR, std::string("run_pipelined: ") + cudaGetErrorString(le));
    return WAE_OK;
}

WAE_API wae_status wae_batch_run_pipelined(wae_batch* b, float* host_out) { return run_pipelined(b, host_out, true); }

WAE_API wae_status wae_batch_sync(wae_batch* b) {
    CUDA_TRY(cudaSetDevice(b->engine->device));
    CUDA_TRY(cudaStreamSynchronize(b->engine->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, b->ev0, b->ev1) == cudaSuccess) b->stats.last_run_ms = ms;
    else cudaGetLastError();
    if (!b->timed.empty()) {
        for (auto& st : b->stages) st.ms = 0.f;
        for (auto& t : b->timed) {
            float v = 0.f;
            if (cudaEventElapsedTime(&v, b->stage_events[t.e0], b->stage_events[t.e1]) == cudaSuccess) b->stages[t.stage].ms += v;
            else cudaGetLastError();
        }
        int best = -1;
        for (size_t i = 0; i < b->stages.size(); i++)
            if (best < 0 || b->stages[i].ms > b->stages[best].ms) best = (int)i;
        if (best >= 0) {
            b->stats.dominant_kernel_ms = b->stages[best].ms;
            const char* nm = kStageNames[b->stages[best].kind];
            std::snprintf(b->stats.dominant_kernel, sizeof(b->stats.dominant_kernel), "%s", nm);
        }
    }
    return WAE_OK;
}

WAE_API wae_status wae_batch_output_device_ptr(wae_batch* b, float** out_dev, uint64_t* out_floats) {
    *out_dev = b->d_out;
    *out_floats = (uint64_t)b->n_graphs * b->channels * b->length;
    return WAE_OK;
}

WAE_API wae_status wae_batch_fetch(wae_batch* b, float* host_out) {
    CUDA_TRY(cudaSetDevice(b->engine->device));
    size_t bytes = (size_t)b->n_graphs * b->channels * b->length * sizeof(float);
    CUDA_TRY(cudaMemcpyAsync(host_out, b->d_out, bytes, cudaMemcpyDeviceToHost, b->engine->stream));
    CUDA_TRY(cudaStreamSynchronize(b->engine->stream));
    return WAE_OK;
}

WAE_API wae_status wae_batch_stage_time(wae_batch* b, uint32_t index, char* name64, float* ms, uint32_t* n_instances) {
    if (index >= b->stages.size()) return fail(WAE_INVALID_ARGUMENT, "stage index out of range");
    const Stage& st = b->stages[index];
    const char* nm = kStageNames[st.kind];
    std::snprintf(name64, 64, "%s", nm);
    *ms = st.ms;
    *n_instances = (uint32_t)st.n;
    return WAE_OK;
}

WAE_API wae_status wae_batch_get_stats(wae_batch* b, wae_batch_stats* out) {
    *out = b->stats;
    return WAE_OK;
}

// ---- one-shot render into a HOST buffer: what `OfflineAudioContext::start_rendering_sync` (src/context/offline.rs:157-185) is for
// a batch of contexts.  Everything a render needs happens inside this call, overlapped:
//   sizing pass (workers, all groups at once)
//   -> H2D of every group's source PCM (copy stream; straight from the graphs' page-locked AudioBuffer memory)
//   -> per group, in order: plan (workers, running ahead) | render (engine stream, waits for the group's PCM) | D2H (copy stream)
//   -> D2H lands in `out` directly when `out` is page-locked, else in one of four page-locked staging slots that worker threads
//      copy out to `out` while the next groups are in flight.
// Device memory comes from the engine's cache (wae_engine::dev_alloc): after the first call of a given shape no cudaMalloc / cudaFree.
// copy with non-temporal stores: the destination (the caller's pageable buffer) is written once and not read here, so its lines are
// not fetched first (a plain memcpy of a few MB stays below glibc's non-temporal threshold and pays a read for every line it writes)
static void copy_streaming(void* dst, const void* src, size_t n) {
    char* d = static_cast<char*>(dst);
    const char* sp = static_cast<const char*>(src);
    const size_t head = std::min(n, (size_t)((64 - (reinterpret_cast<uintptr_t>(d) & 63)) & 63));
    if (head) std::memcpy(d, sp, head);
    d += head; sp += head; n -= head;
    if ((reinterpret_cast<uintptr_t>(sp) & 15) == 0) {
        for (; n >= 64; n -= 64, d += 64, sp += 64) {
            const __m128i a = _mm_load_si128(reinterpret_cast<const __m128i*>(sp)), b2 = _mm_load_si128(reinterpret_cast<const __m128i*>(sp + 16));
            const __m128i c = _mm_load_si128(reinterpret_cast<const __m128i*>(sp + 32)), e = _mm_load_si128(reinterpret_cast<const __m128i*>(sp + 48));
            _mm_stream_si128(reinterpret_cast<__m128i*>(d), a);
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + 16), b2);
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + 48), e);
        }
    } else {
        for (; n >= 64; n -= 64, d += 64, sp += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(sp)), b2 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(sp + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(sp + 32)), e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(sp + 48));
            _mm_stream_si128(reinterpret_cast<__m128i*>(d), a);
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + 16), b2);
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + 48), e);
        }
    }
    _mm_sfence();
    if (n) std::memcpy(d, sp, n);
}

static wae_status render_oneshot_host(wae_engine* eng, wae_graph* const* graphs, uint32_t n_graphs, float* out) {
    if (!out) return fail(WAE_INVALID_ARGUMENT, "null output buffer");
    PrepState ps;
    wae_batch* b = nullptr;
    wae_status st = prep_begin(eng, graphs, n_graphs, nullptr, &b, ps, true);
    if (st != WAE_OK) return st;
    const int n_groups = (int)b->groups.size();
    const size_t per_graph = (size_t)b->channels * b->length;
    cudaStream_t s = eng->stream;
    WorkerPool* pool = eng->workers();
    // ---- everything below must run to its end before the batch can be destroyed: `pending` counts worker tasks in flight
    std::mutex mu;
    std::condition_variable cv;
    int pending = 0;
    std::vector<char> planned(n_groups, 0);
    std::vector<GroupPlan> gps(n_groups);
    auto task_done = [&] {
        std::lock_guard<std::mutex> lk(mu);
        pending--;
        cv.notify_all();
    };
    auto drain = [&] {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return pending == 0; });
    };
    wae_status result = WAE_OK;
    std::string result_msg;
    auto set_fail = [&](wae_status code, const std::string& msg) {
        if (result == WAE_OK) {
            result = code;
            result_msg = msg;
        }
    };
    // 1. the source PCM of all groups, in group order, on the H2D stream
    for (int k = 0; k < n_groups && result == WAE_OK; k++) {
        wae_batch::Group& grp = b->groups[k];
        if (!grp.src_floats) continue;
        if (enqueue_source_copies(b, grp, eng->s_h2d) != WAE_OK || cudaEventRecord(grp.ev_h2d, eng->s_h2d) != cudaSuccess)
            set_fail(WAE_CUDA_ERROR, std::string("one-shot render: H2D of the source PCM failed: ") + wae_last_error());
    }
    // 2. plans, running ahead of the render on the workers
    if (result == WAE_OK) {
        std::lock_guard<std::mutex> lk(mu);
        pending += n_groups;
    }
    if (result == WAE_OK)
        for (int k = 0; k < n_groups; k++)
            pool->submit([&, k] {
                prep_plan_group(b, graphs, k, ps, gps[k]);
                {
                    std::lock_guard<std::mutex> lk(mu);
                    planned[k] = 1;
                }
                task_done();
            });
    // 3. where the rendered PCM lands
    bool out_pinned = false;
    {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, out) == cudaSuccess) out_pinned = attr.type == cudaMemoryTypeHost;
        else cudaGetLastError();
    }
    size_t max_group_bytes = 0;
    for (auto& grp : b->groups) max_group_bytes = std::max(max_group_bytes, (size_t)(grp.g1 - grp.g0) * per_graph * sizeof(float));
    // Pageable `out`, default: whole-group page-locked staging slots, copied out in parts by the workers (below).  WAE_STAGE_RING=1 (an
    // experiment that lost, kept for the record: profiles/README.md r2_x): the PCM comes down in PIECES of a couple of MB through a small
    // ring of page-locked slots meant to stay in the last-level cache (inbound DMA writes allocate there), each piece copied out as soon
    // as it has landed.  Two ranks on one socket: 414 - 1117 ms per call against 184 ms with the group slots — a piece pays a blocking
    // event wait and two thread wake-ups, and 2 MB is not enough work to hide them.
    static const bool use_ring = [] { const char* e = getenv("WAE_STAGE_RING"); return e && atoi(e) != 0; }();
    static const size_t piece_bytes = [] { const char* e = getenv("WAE_STAGE_PIECE_KB"); long kb = e ? atol(e) : 2048; return (size_t)std::max(64l, std::min(65536l, kb)) * 1024; }();
    static const int ring_slots = [] { const char* e = getenv("WAE_STAGE_SLOTS"); int n = e ? atoi(e) : 8; return std::max(2, std::min(64, n)); }();
    const bool ring = !out_pinned && use_ring;
    if (result == WAE_OK && ring && !eng->ensure_ring(piece_bytes * (size_t)ring_slots)) set_fail(WAE_OUT_OF_MEMORY, "out of memory (page-locked staging ring of the rendered PCM)");
    if (result == WAE_OK && !out_pinned && !ring && !eng->ensure_stage(max_group_bytes)) set_fail(WAE_OUT_OF_MEMORY, "out of memory (page-locked staging of the rendered PCM)");
    // ring state (guarded by `mu`): slot i is free again once its copy-out is done; groups are handed to the pump thread in order
    std::vector<char> ring_busy(ring ? ring_slots : 0, 0);
    std::vector<cudaEvent_t> ring_ev(ring ? ring_slots : 0, nullptr);
    for (auto& e : ring_ev)
        if (result == WAE_OK && cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess) set_fail(WAE_CUDA_ERROR, "cudaEventCreate failed");
    int issued_groups = 0;       // groups whose render has been launched and whose ev_done is recorded
    bool pump_abort = false;     // the main thread gave up: no more groups will come
    std::string pump_error;
    std::thread pump;
    if (result == WAE_OK && ring)
        pump = std::thread([&] {
            cudaSetDevice(eng->device);
            size_t piece_no = 0;
            for (int k = 0; k < n_groups; k++) {
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return issued_groups > k || pump_abort; });
                    if (issued_groups <= k) return;
                }
                wae_batch::Group& grp = b->groups[k];
                if (cudaStreamWaitEvent(eng->s_d2h, grp.ev_done, 0) != cudaSuccess) {
                    std::lock_guard<std::mutex> lk(mu);
                    pump_error = "cudaStreamWaitEvent failed";
                    return;
                }
                const size_t off = (size_t)grp.g0 * per_graph * sizeof(float), bytes = (size_t)(grp.g1 - grp.g0) * per_graph * sizeof(float);
                for (size_t a0 = 0; a0 < bytes; a0 += piece_bytes, piece_no++) {
                    const size_t nb = std::min(piece_bytes, bytes - a0);
                    const int slot = (int)(piece_no % (size_t)ring_slots);
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [&] { return !ring_busy[slot]; });
                        ring_busy[slot] = 1;
                        pending++;
                    }
                    char* stage = eng->h_ring + (size_t)slot * piece_bytes;
                    if (cudaMemcpyAsync(stage, (const char*)b->d_out + off + a0, nb, cudaMemcpyDeviceToHost, eng->s_d2h) != cudaSuccess ||
                        cudaEventRecord(ring_ev[slot], eng->s_d2h) != cudaSuccess) {
                        std::lock_guard<std::mutex> lk(mu);
                        pump_error = "D2H failed";
                        ring_busy[slot] = 0;
                        pending--;
                        cv.notify_all();
                        return;
                    }
                    char* dst = (char*)out + off + a0;
                    pool->submit([&, slot, stage, dst, nb] {
                        cudaEventSynchronize(ring_ev[slot]);
                        copy_streaming(dst, stage, nb);
                        {
                            std::lock_guard<std::mutex> lk(mu);
                            ring_busy[slot] = 0;
                        }
                        task_done();  // (notifies: the pump may be waiting for this slot)
                    });
                }
            }
        });
    constexpr int SLOTS = wae_engine::kStageSlots;
    bool slot_busy[SLOTS] = {false, false, false, false};
    std::vector<cudaEvent_t> ev_copy((out_pinned || ring) ? 0 : n_groups, nullptr);
    std::vector<std::unique_ptr<std::atomic<int>>> parts_left;
    for (auto& e : ev_copy)
        if (result == WAE_OK && cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess)
            set_fail(WAE_CUDA_ERROR, "cudaEventCreate failed");
    static const int env_parts = [] { const char* e = getenv("WAE_COPY_PARTS"); return e ? atoi(e) : 0; }();  // (tuning)
    const int copy_parts = env_parts > 0 ? std::min(env_parts, 32) : std::max(1, std::min(8, pool->size() / 2));
    if (result == WAE_OK) {
        if (cudaEventCreate(&b->ev0) != cudaSuccess || cudaEventCreate(&b->ev1) != cudaSuccess || cudaEventRecord(b->ev0, s) != cudaSuccess)
            set_fail(WAE_CUDA_ERROR, "cudaEventCreate failed");
    }
    // 4. group by group: wait for its plan, render, copy back
    for (int k = 0; k < n_groups && result == WAE_OK; k++) {
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return planned[k] != 0; });
        }
        if (gps[k].code != WAE_OK) {
            set_fail(gps[k].code, gps[k].error);
            break;
        }
        prep_append_group(b, k, ps, gps[k]);
        wae_batch::Group& grp = b->groups[k];
        if (grp.src_floats && cudaStreamWaitEvent(s, grp.ev_h2d, 0) != cudaSuccess) {
            set_fail(WAE_CUDA_ERROR, "cudaStreamWaitEvent failed");
            break;
        }
        if (run_group(b, grp) != WAE_OK) {
            set_fail(WAE_CUDA_ERROR, wae_last_error());
            break;
        }
        const size_t off = (size_t)grp.g0 * per_graph, bytes = (size_t)(grp.g1 - grp.g0) * per_graph * sizeof(float);
        // (ring: the pump thread makes the copy stream wait, in group order — a wait queued from here could land between the pieces of
        // the group before)
        if (cudaEventRecord(grp.ev_done, s) != cudaSuccess || (!ring && cudaStreamWaitEvent(eng->s_d2h, grp.ev_done, 0) != cudaSuccess)) {
            set_fail(WAE_CUDA_ERROR, "event record / wait failed");
            break;
        }
        if (out_pinned) {
            if (cudaMemcpyAsync(out + off, b->d_out + off, bytes, cudaMemcpyDeviceToHost, eng->s_d2h) != cudaSuccess) set_fail(WAE_CUDA_ERROR, "D2H failed");
            continue;
        }
        if (ring) {  // the pump thread brings this group down piece by piece
            std::lock_guard<std::mutex> lk(mu);
            issued_groups = k + 1;
            cv.notify_all();
            continue;
        }
        const int slot = k % SLOTS;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return !slot_busy[slot]; });
            slot_busy[slot] = true;
            pending += copy_parts;
        }
        if (cudaMemcpyAsync(eng->h_stage[slot], b->d_out + off, bytes, cudaMemcpyDeviceToHost, eng->s_d2h) != cudaSuccess ||
            cudaEventRecord(ev_copy[k], eng->s_d2h) != cudaSuccess) {
            set_fail(WAE_CUDA_ERROR, "D2H failed");
            std::lock_guard<std::mutex> lk(mu);
            pending -= copy_parts;
            slot_busy[slot] = false;
            break;
        }
        parts_left.emplace_back(new std::atomic<int>(copy_parts));
        std::atomic<int>* left = parts_left.back().get();
        for (int part = 0; part < copy_parts; part++)
            pool->submit([&, k, slot, part, left, off, bytes] {
                cudaEventSynchronize(ev_copy[k]);
                const size_t chunk = (bytes / copy_parts + 63) / 64 * 64;
                const size_t a0 = std::min(bytes, (size_t)part * chunk), a1 = std::min(bytes, a0 + chunk);
                // (non-temporal stores: a 15 MB part is far below glibc's non-temporal threshold — 3/4 of a 260 MB L3 — and a plain memcpy
                // would fetch every destination line before overwriting it; WAE_STAGE_NT=0: memcpy)
                static const bool nt = [] { const char* e = getenv("WAE_STAGE_NT"); return !e || atoi(e) != 0; }();
                if (a1 > a0) {
                    if (nt) copy_streaming((char*)(out + off) + a0, (const char*)eng->h_stage[slot] + a0, a1 - a0);
                    else std::memcpy((char*)(out + off) + a0, (const char*)eng->h_stage[slot] + a0, a1 - a0);
                }
                if (left->fetch_sub(1) == 1) {
                    std::lock_guard<std::mutex> lk(mu);
                    slot_busy[slot] = false;
                }
                task_done();
            });
    }
    if (result == WAE_OK && cudaEventRecord(b->ev1, s) != cudaSuccess) set_fail(WAE_CUDA_ERROR, "cudaEventRecord failed");
    if (pump.joinable()) {
        {
            std::lock_guard<std::mutex> lk(mu);
            if (issued_groups < n_groups) pump_abort = true;  // (a failure above: the groups that were launched are still brought down)
            cv.notify_all();
        }
        pump.join();
        if (!pump_error.empty()) set_fail(WAE_CUDA_ERROR, "one-shot render: " + pump_error);
    }
    drain();  // plans and copy-outs
    if (cudaStreamSynchronize(eng->s_d2h) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess || cudaStreamSynchronize(eng->s_h2d) != cudaSuccess)
        set_fail(WAE_CUDA_ERROR, std::string("one-shot render: ") + cudaGetErrorString(cudaGetLastError()));
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) set_fail(WAE_CUDA_ERROR, std::string("one-shot render: ") + cudaGetErrorString(le));
    for (auto& e : ev_copy)
        if (e) cudaEventDestroy(e);
    for (auto& e : ring_ev)
        if (e) cudaEventDestroy(e);
    wae_batch_destroy(b);
    if (result != WAE_OK) return fail(result, result_msg);
    return WAE_OK;
}

// Page-locked host memory for callers that keep their output (or input) buffers around: D2H lands in such a buffer directly instead of
// going through the staging slots.  wae_host_register page-locks memory the caller allocated itself (it must stay allocated until
// wae_host_unregister); both are thin wrappers so that a binding needs no CUDA of its own.
WAE_API wae_status wae_host_alloc(wae_engine* eng, uint64_t bytes, void** out) {
    if (!eng || !out || bytes == 0) return fail(WAE_INVALID_ARGUMENT, "null engine / out pointer or zero size");
    CUDA_TRY(cudaSetDevice(eng->device));
    if (cudaHostAlloc(out, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return fail(WAE_OUT_OF_MEMORY, "out of page-locked host memory");
    }
    return WAE_OK;
}
WAE_API wae_status wae_host_free(wae_engine* eng, void* p) {
    if (!eng) return fail(WAE_INVALID_ARGUMENT, "null engine");
    if (p) CUDA_TRY(cudaFreeHost(p));
    return WAE_OK;
}
WAE_API wae_status wae_host_register(wae_engine* eng, void* p, uint64_t bytes) {
    if (!eng || !p || bytes == 0) return fail(WAE_INVALID_ARGUMENT, "null engine / pointer or zero size");
    CUDA_TRY(cudaSetDevice(eng->device));
    if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) {
        const std::string msg = std::string("cudaHostRegister: ") + cudaGetErrorString(cudaGetLastError());
        return fail(WAE_CUDA_ERROR, msg);
    }
    return WAE_OK;
}
WAE_API wae_status wae_host_unregister(wae_engine* eng, void* p) {
    if (!eng || !p) return fail(WAE_INVALID_ARGUMENT, "null engine / pointer");
    CUDA_TRY(cudaHostUnregister(p));
    return WAE_OK;
}

WAE_API wae_status wae_selftest_conv_fft(float* data, uint32_t mode) {
    if (!data || mode > 3) return fail(WAE_INVALID_ARGUMENT, "wae_selftest_conv_fft: null data or mode > 3");
    conv_fft_selftest(data, (int)mode);
    return WAE_OK;
}

WAE_API wae_status wae_render_batch(wae_engine* eng, wae_graph* const* graphs, uint32_t n_graphs, float* out, uint32_t flags) {
    if (!(flags & WAE_RENDER_OUT_DEVICE)) return render_oneshot_host(eng, graphs, n_graphs, out);
    wae_batch* b = nullptr;
    wae_status st = wae_batch_prepare(eng, graphs, n_graphs, &b);
    if (st != WAE_OK) return st;
    st = wae_batch_run(b);
    if (st == WAE_OK) st = wae_batch_sync(b);
    if (st == WAE_OK) {
        size_t bytes = (size_t)b->n_graphs * b->channels * b->length * sizeof(float);
        cudaError_t e = cudaMemcpyAsync(out, b->d_out, bytes, cudaMemcpyDeviceToDevice, eng->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(eng->stream);
        if (e != cudaSuccess) st = fail(WAE_CUDA_ERROR, cudaGetErrorString(e));
    }
    std::string saved = wae_last_error();
    wae_batch_destroy(b);
    if (st != WAE_OK) set_error(saved);
    return st;
}

// AnalyserNode read-out: get_float_time_domain_data (src/analysis.rs:261-264, ring read :114-127)
static const AnalyserRec* find_analyser(wae_batch* b, uint32_t gi, wae_node_id node) {
    for (auto& a : b->analysers)
        if (a.graph_index == gi && a.node == node) return &a;
    return nullptr;
}

WAE_API wae_status wae_analyser_get_float_time_domain_data(wae_batch* b, uint32_t graph_index, wae_node_id node, float* out, uint32_t len) {
    const AnalyserRec* a = find_analyser(b, graph_index, node);
    if (!a) return fail(WAE_INVALID_ARGUMENT, "not an analyser of this batch");
    const uint32_t RING = 32768 + 128;
    std::vector<float> ring(RING);
    CUDA_TRY(cudaSetDevice(b->engine->device));
    CUDA_TRY(cudaMemcpy(ring.data(), a->d_ring, RING * sizeof(float), cudaMemcpyDeviceToHost));
    uint32_t n = std::min(len, a->fft_size);
    uint64_t write_index = (uint64_t)b->lq % RING;
    for (uint32_t i = 0; i < n; i++) out[i] = ring[(RING + write_index - n + i) % RING];
    return WAE_OK;
}

WAE_API wae_status wae_analyser_get_float_frequency_data(wae_batch* b, uint32_t graph_index, wae_node_id node, float* out, uint32_t len) {
    AnalyserRec* a = const_cast<AnalyserRec*>(find_analyser(b, graph_index, node));
    if (!a) return fail(WAE_INVALID_ARGUMENT, "not an analyser of this batch");
    CUDA_TRY(cudaSetDevice(b->engine->device));
    const uint32_t RING = 32768 + 128;
    const uint32_t bins = a->fft_size / 2;
    if (!a->computed) {  // one FFT per distinct current_time (analysis.rs:353-361): the read-out happens after the render
        launch_analyser_fft(a->d_ring, (uint32_t)((uint64_t)b->lq % RING), (int)a->fft_size, (float)a->smoothing, a->d_last_fft, a->d_db,
                            b->engine->stream);
        a->computed = true;
    }
    uint32_t n = std::min(len, bins);
    CUDA_TRY(cudaMemcpyAsync(out, a->d_db, n * sizeof(float), cudaMemcpyDeviceToHost, b->engine->stream));
    CUDA_TRY(cudaStreamSynchronize(b->engine->stream));
    return WAE_OK;
}

// Analyser::get_byte_time_domain_data (src/analysis.rs:266-276): 128 (1 + x) clamped to a byte
WAE_API wae_status wae_analyser_get_byte_time_domain_data(wae_batch* b, uint32_t graph_index, wae_node_id node, uint8_t* out, uint32_t len) {
    std::vector<float> tmp(len, 0.f);
    wae_status st = wae_analyser_get_float_time_domain_data(b, graph_index, node, tmp.data(), len);
    if (st != WAE_OK) return st;
    const AnalyserRec* a = find_analyser(b, graph_index, node);
    const uint32_t n = std::min(len, a->fft_size);
    for (uint32_t i = 0; i < n; i++) {
        float scaled = 128.f * (1.f + tmp[i]);
        scaled = scaled < 0.f ? 0.f : (scaled > 255.f ? 255.f : scaled);
        out[i] = (uint8_t)scaled;
    }
    return WAE_OK;
}

// Analyser::get_byte_frequency_data (src/analysis.rs:371-401): dB scaled into [minDecibels, maxDecibels] -> 0..255
WAE_API wae_status wae_analyser_get_byte_frequency_data(wae_batch* b, uint32_t graph_index, wae_node_id node, uint8_t* out, uint32_t len) {
    const AnalyserRec* a = find_analyser(b, graph_index, node);
    if (!a) return fail(WAE_INVALID_ARGUMENT, "not an analyser of this batch");
    const uint32_t n = std::min(len, a->fft_size / 2);
    std::vector<float> db(n, 0.f);
    wae_status st = wae_analyser_get_float_frequency_data(b, graph_index, node, db.data(), n);
    if (st != WAE_OK) return st;
    const float mn = (float)a->min_db, mx = (float)a->max_db;
    for (uint32_t i = 0; i < n; i++) {
        float scaled = 255.f / (mx - mn) * (db[i] - mn);
        scaled = !(scaled > 0.f) ? 0.f : (scaled > 255.f ? 255.f : scaled);  // -inf dB (silence) and NaN -> 0
        out[i] = (uint8_t)scaled;
    }
    return WAE_OK;
}

// DynamicsCompressorNode::reduction (src/node/dynamics_compressor.rs:204-206,448): the reduction (dB) of the last frame
WAE_API wae_status wae_compressor_reduction(wae_batch* b, uint32_t graph_index, wae_node_id node, float* out) {
    if (!b || !out) return fail(WAE_INVALID_ARGUMENT, "null argument");
    for (auto& c : b->compressors)
        if (c.graph == graph_index && c.node == node) {
            CUDA_TRY(cudaSetDevice(b->engine->device));
            CUDA_TRY(cudaMemcpyAsync(out, c.d_state + 1, sizeof(float), cudaMemcpyDeviceToHost, b->engine->stream));
            CUDA_TRY(cudaStreamSynchronize(b->engine->stream));
            return WAE_OK;
        }
    return fail(WAE_INVALID_ARGUMENT, "not a dynamics compressor of this batch");
}

// AudioBuffer::resample on the GPU (src/buffer.rs:311-363): `in` / `out` are host pointers, out_cap >= ceil(len * to/from)
WAE_API wae_status wae_resample_linear(wae_engine* eng, const float* in, uint64_t len, float from_rate, float to_rate, float* out,
                                       uint64_t out_cap, uint64_t* out_len) {
    if (!eng || !in || !out || !out_len) return fail(WAE_INVALID_ARGUMENT, "null argument");
    if (std::fabs(from_rate - to_rate) <= 0.1f || len == 0) {  // "very similar sample rate: do not resample"
        if (out_cap < len) return fail(WAE_INVALID_ARGUMENT, "output capacity too small");
        std::memcpy(out, in, len * sizeof(float));
        *out_len = len;
        return WAE_OK;
    }
    uint64_t target = (uint64_t)std::ceil((double)len * ((double)to_rate / (double)from_rate));
    if (out_cap < target) return fail(WAE_INVALID_ARGUMENT, "output capacity too small");
    CUDA_TRY(cudaSetDevice(eng->device));
    float *d_in = nullptr, *d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_in, len * sizeof(float)));
    if (cudaMalloc(&d_out, target * sizeof(float)) != cudaSuccess) {
        cudaFree(d_in);
        return fail(WAE_OUT_OF_MEMORY, "out of device memory (resample)");
    }
    cudaMemcpyAsync(d_in, in, len * sizeof(float), cudaMemcpyHostToDevice, eng->stream);
    launch_resample_linear(d_in, (int64_t)len, d_out, (int64_t)target, eng->stream);
    cudaMemcpyAsync(out, d_out, target * sizeof(float), cudaMemcpyDeviceToHost, eng->stream);
    cudaError_t e = cudaStreamSynchronize(eng->stream);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(WAE_CUDA_ERROR, cudaGetErrorString(e));
    *out_len = target;
    return WAE_OK;
}

}  // extern "C"
