// fmb_engine.cu -- the one-call end-to-end path (fmc::Search{index, queries, ...}() of the reference, search/search.h:47-75):
// host queries in, located rows out.  fmb_search_and_locate, its 2-bit packed variant and the multi-GPU variant share one engine.
//
// Every index owns an engine: a few persistent host threads, each with its own CUDA stream.  A call cuts its query range into
// chunks; a worker takes the next chunk and drives upload -> search -> locate -> download of that chunk on its stream, so the H2D
// copy of one chunk, the kernels of another and the D2H copy of a third overlap.  Rows are written in CHUNK ORDER (a chunk learns
// its offset from its predecessor's row count), i.e. grouped by ascending ranges of qidx, whatever the completion order was.
// Multi GPU: one replica of the index per device, the query range is split into contiguous shards, one per replica, every replica's
// engine works on its shard at the same time and writes to its own segment of the output; nothing is exchanged between GPUs.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "fmb_host.hpp"

using namespace fmb;

namespace fmb {

struct Job {
    const fmb_index* ix = nullptr;
    // queries [q_begin, q_end) of the caller's batch: byte symbols, or 2-bit packed words with an exception list
    const uint8_t* symbols = nullptr;
    const uint32_t* packed = nullptr;
    const uint64_t* exc_pos = nullptr;
    const uint8_t* exc_sym = nullptr;
    uint64_t n_exc = 0;
    const uint64_t* offsets = nullptr;
    uint64_t q_begin = 0, q_end = 0;
    int edit = 0;
    uint32_t n_searches = 0, n_parts = 0;
    const uint32_t *pi = nullptr, *l = nullptr, *u = nullptr, *partition = nullptr;
    fmb_loc32* out = nullptr;
    uint64_t capacity = 0;
    // progress
    uint64_t chunk = 0, n_chunks = 0;
    std::atomic<uint64_t> next_chunk{0};
    std::vector<uint64_t> chain_end;                 // rows of chunks [0, c]
    std::vector<std::atomic<int>> chain_ready;
    std::atomic<int> err{FMB_OK};
    std::mutex mu;
    std::string err_msg;
    fmb_stats total{};
    // completion
    int active = 0;
    bool done = false;
};

struct Engine {
    const fmb_index* ix;
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    Job* job = nullptr;
    uint64_t generation = 0;
    bool stop = false;
    std::mutex call_mu;                              // one call at a time per index

    Engine(const fmb_index* ix_, int n_workers) : ix(ix_) {
        for (int i = 0; i < n_workers; ++i) threads.emplace_back([this] { worker(); });
    }
    ~Engine() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_job.notify_all();
        for (auto& t : threads) t.join();
    }
    void start(Job* j) {
        {
            std::lock_guard<std::mutex> lk(mu);
            j->active = (int)threads.size();
            j->done = false;
            job = j;
            ++generation;
        }
        cv_job.notify_all();
    }
    void wait(Job* j) {
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return j->done; });
        job = nullptr;
    }
    void worker();
    void run_chunks(Job& j, cudaStream_t st);
};

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void Engine::worker() {
    cudaSetDevice(ix->device);
    cudaStream_t st = nullptr;
    const bool have_stream = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
    uint64_t seen = 0;
    for (;;) {
        Job* j;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_job.wait(lk, [&] { return stop || generation != seen; });
            if (stop) break;
            seen = generation;
            j = job;
        }
        if (j) {
            if (!have_stream) {
                std::lock_guard<std::mutex> lk(j->mu);
                if (j->err == FMB_OK) { j->err = FMB_ECUDA; j->err_msg = "cudaStreamCreate failed"; }
            } else {
                tls_stream_override = st;
                run_chunks(*j, st);
                tls_stream_override = nullptr;
            }
            std::lock_guard<std::mutex> lk(mu);
            if (--j->active == 0) {
                j->done = true;
                cv_done.notify_all();
            }
        }
    }
    if (have_stream) cudaStreamDestroy(st);
}

void Engine::run_chunks(Job& j, cudaStream_t st) {
    static const bool trace = getenv("FMB_TRACE") != nullptr;
    double t_up = 0, t_search = 0, t_loc = 0, t_wait = 0, t_down = 0;
    const double t_begin = now_ms();
    fmb_stats mine{};
    auto fail = [&](int rc, const char* msg) {
        std::lock_guard<std::mutex> lk(j.mu);
        if (j.err == FMB_OK || (j.err == FMB_EOVERFLOW && rc != FMB_EOVERFLOW)) { j.err = rc; j.err_msg = msg; }
    };
    // a chunk publishes the number of rows of all chunks up to itself, whatever happened to it (successors never hang)
    auto publish = [&](uint64_t c, uint64_t cnt, uint64_t& off) {
        double t0 = now_ms();
        if (c > 0) while (!j.chain_ready[c - 1].load(std::memory_order_acquire)) std::this_thread::yield();
        off = c ? j.chain_end[c - 1] : 0;
        j.chain_end[c] = off + cnt;
        j.chain_ready[c].store(1, std::memory_order_release);
        t_wait += now_ms() - t0;
    };
    for (;;) {
        const uint64_t c = j.next_chunk.fetch_add(1);
        if (c >= j.n_chunks) break;
        uint64_t off = 0;
        const int e0 = j.err;
        if (e0 != FMB_OK && e0 != FMB_EOVERFLOW) { publish(c, 0, off); continue; }
        const uint64_t b = j.q_begin + c * j.chunk, e = std::min(j.q_end, b + j.chunk);
        fmb_queries* q = nullptr;
        double t0 = now_ms();
        int rc = j.packed ? fmb_queries_upload_packed(&q, j.ix, j.packed, j.offsets + b, e - b, j.exc_pos, j.exc_sym, j.n_exc)
                          : fmb_queries_upload(&q, j.ix, j.symbols, j.offsets + b, e - b);
        if (rc) { fail(rc, fmb_last_error()); publish(c, 0, off); continue; }
        q->qidx_base = b;
        mine.h2d_bytes += q->h2d_bytes;
        double t1 = now_ms();
        fmb_results* hits = nullptr;
        rc = j.n_searches ? fmb_search_scheme(j.ix, q, j.edit, j.n_searches, j.n_parts, j.pi, j.l, j.u, j.partition, &hits) : fmb_search_exact(j.ix, q, &hits);
        fmb_queries_destroy(q);
        if (rc) { fail(rc, fmb_last_error()); publish(c, 0, off); continue; }
        double t2 = now_ms();
        fmb_results* locs = nullptr;
        rc = fmb_locate(j.ix, hits, &locs);
        double t3 = now_ms();
        t_up += t1 - t0; t_search += t2 - t1; t_loc += t3 - t2;
        mine.extensions += hits->stats.extensions;
        mine.occ_lookups += hits->stats.occ_lookups;
        mine.line_requests += hits->stats.line_requests;
        mine.kernel_ms += hits->stats.kernel_ms;
        mine.main_kernel_ms += hits->stats.main_kernel_ms;
        mine.frontier_peak = std::max(mine.frontier_peak, hits->stats.frontier_peak);
        fmb_results_destroy(hits);
        if (rc) { fail(rc, fmb_last_error()); publish(c, 0, off); continue; }
        mine.lf_steps += locs->stats.lf_steps;
        mine.occ_lookups += locs->stats.occ_lookups;
        mine.kernel_ms += locs->stats.kernel_ms;
        const uint64_t cnt = locs->count;
        publish(c, cnt, off);
        if (off + cnt > j.capacity) {
            fmb_results_destroy(locs);
            fail(FMB_EOVERFLOW, "output capacity too small");
            continue;                        // keep counting so that the caller learns the size that is needed
        }
        double t4 = now_ms();
        cudaError_t ce = cudaSuccess;
        mine.d2h_bytes += cnt * sizeof(fmb_loc32);
        if (cnt) ce = cudaMemcpyAsync(j.out + off, locs->locs.p, cnt * sizeof(fmb_loc32), cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        t_down += now_ms() - t4;
        fmb_results_destroy(locs);
        if (ce != cudaSuccess) fail(FMB_ECUDA, (std::string("D2H of located rows: ") + cudaGetErrorString(ce)).c_str());
    }
    std::lock_guard<std::mutex> lk(j.mu);
    if (trace)
        fprintf(stderr, "[fmb trace] worker done after %.2f ms: upload %.2f search %.2f locate %.2f wait %.2f download %.2f\n", now_ms() - t_begin, t_up,
                t_search, t_loc, t_wait, t_down);
    j.total.extensions += mine.extensions;
    j.total.occ_lookups += mine.occ_lookups;
    j.total.line_requests += mine.line_requests;
    j.total.lf_steps += mine.lf_steps;
    j.total.h2d_bytes += mine.h2d_bytes;
    j.total.d2h_bytes += mine.d2h_bytes;
    j.total.kernel_ms += mine.kernel_ms;
    j.total.main_kernel_ms += mine.main_kernel_ms;
    j.total.frontier_peak = std::max(j.total.frontier_peak, mine.frontier_peak);
}

static Engine* engine_of(const fmb_index* ix) {
    std::lock_guard<std::mutex> lk(ix->engine_mu);
    if (!ix->engine) {
        static const int env_threads = getenv("FMB_E2E_THREADS") ? atoi(getenv("FMB_E2E_THREADS")) : 0;
        ix->engine = new Engine(ix, std::max(1, env_threads ? env_threads : 6));
    }
    return static_cast<Engine*>(ix->engine);
}

void engine_destroy(fmb_index* ix) {
    std::lock_guard<std::mutex> lk(ix->engine_mu);
    delete static_cast<Engine*>(ix->engine);
    ix->engine = nullptr;
}

static void prepare(Job& j) {
    static const int env_chunk = getenv("FMB_E2E_CHUNK_LOG2") ? atoi(getenv("FMB_E2E_CHUNK_LOG2")) : 0;
    // measured with 10 M reads: exact search (PCIe bound) 2^19 x 6 threads beats 2^20 x 3 and 2^18 x 12; the k-error searches run several
    // kernels per chunk and want larger launches: 2^20 (k = 1 edit 270 -> 312 M reads/s, k = 2 edit 111 -> 121)
    j.chunk = env_chunk ? (uint64_t(1) << env_chunk) : (j.n_searches ? (1u << 20) : (1u << 19));
    const uint64_t nq = j.q_end - j.q_begin;
    j.n_chunks = (nq + j.chunk - 1) / j.chunk;
    j.chain_end.assign(j.n_chunks, 0);
    j.chain_ready = std::vector<std::atomic<int>>(j.n_chunks);
    for (auto& r : j.chain_ready) r.store(0);
}

// runs the jobs (one per replica) at the same time; *n_out[g] = rows of job g (or the capacity it would have needed)
static int run_jobs(std::vector<Job>& jobs, uint64_t* n_out, fmb_stats* stats) {
    std::vector<Engine*> engines;
    for (auto& j : jobs) engines.push_back(engine_of(j.ix));
    // one call at a time per index; the locks are taken in address order so that concurrent multi-replica calls cannot deadlock
    std::vector<Engine*> order = engines;
    std::sort(order.begin(), order.end());
    std::vector<std::unique_lock<std::mutex>> locks;
    for (auto* e : order) locks.emplace_back(e->call_mu);
    for (size_t g = 0; g < jobs.size(); ++g) {
        prepare(jobs[g]);
        if (jobs[g].n_chunks) engines[g]->start(&jobs[g]);
    }
    for (size_t g = 0; g < jobs.size(); ++g)
        if (jobs[g].n_chunks) engines[g]->wait(&jobs[g]);
    int rc = FMB_OK;
    std::string msg;
    fmb_stats total{};
    for (size_t g = 0; g < jobs.size(); ++g) {
        Job& j = jobs[g];
        n_out[g] = j.n_chunks ? j.chain_end[j.n_chunks - 1] : 0;
        total.extensions += j.total.extensions;
        total.occ_lookups += j.total.occ_lookups;
        total.line_requests += j.total.line_requests;
        total.lf_steps += j.total.lf_steps;
        total.h2d_bytes += j.total.h2d_bytes;
        total.d2h_bytes += j.total.d2h_bytes;
        total.kernel_ms += j.total.kernel_ms;
        total.main_kernel_ms += j.total.main_kernel_ms;
        total.frontier_peak = std::max(total.frontier_peak, j.total.frontier_peak);
        if (j.err != FMB_OK && (rc == FMB_OK || rc == FMB_EOVERFLOW)) {
            rc = j.err;
            msg = j.err == FMB_EOVERFLOW ? "output capacity " + std::to_string(j.capacity) + " too small, " + std::to_string(n_out[g]) + " rows found"
                                         : j.err_msg;
        }
    }
    if (stats) *stats = total;
    if (rc != FMB_OK) set_error("%s", msg.c_str());
    return rc;
}

static int check_scheme_args(uint32_t n_searches, const uint32_t* pi, const uint32_t* l, const uint32_t* u, const uint32_t* partition) {
    if (n_searches && (!pi || !l || !u || !partition)) { set_error("NULL scheme argument"); return FMB_EINVAL; }
    return FMB_OK;
}

}  // namespace fmb

extern "C" {

int fmb_search_and_locate(const fmb_index* ix, const uint8_t* symbols, const uint64_t* offsets, uint64_t nq, int edit,
                          uint32_t n_searches, uint32_t n_parts, const uint32_t* pi, const uint32_t* l, const uint32_t* u,
                          const uint32_t* partition, fmb_loc32* out, uint64_t capacity, uint64_t* n_out, fmb_stats* stats) {
    if (!ix || !offsets || !n_out || (capacity && !out)) { set_error("NULL argument"); return FMB_EINVAL; }
    *n_out = 0;
    if (stats) *stats = fmb_stats{};
    FMB_TRY(check_scheme_args(n_searches, pi, l, u, partition));
    if (nq == 0) return FMB_OK;
    std::vector<Job> jobs(1);
    Job& j = jobs[0];
    j.ix = ix; j.symbols = symbols; j.offsets = offsets; j.q_begin = 0; j.q_end = nq;
    j.edit = edit; j.n_searches = n_searches; j.n_parts = n_parts; j.pi = pi; j.l = l; j.u = u; j.partition = partition;
    j.out = out; j.capacity = capacity;
    return run_jobs(jobs, n_out, stats);
}

int fmb_search_and_locate_packed(const fmb_index* ix, const uint32_t* packed, const uint64_t* offsets, uint64_t nq,
                                 const uint64_t* exc_pos, const uint8_t* exc_sym, uint64_t n_exc, int edit,
                                 uint32_t n_searches, uint32_t n_parts, const uint32_t* pi, const uint32_t* l, const uint32_t* u,
                                 const uint32_t* partition, fmb_loc32* out, uint64_t capacity, uint64_t* n_out, fmb_stats* stats) {
    if (!ix || !offsets || !n_out || (capacity && !out) || (n_exc && (!exc_pos || !exc_sym))) { set_error("NULL argument"); return FMB_EINVAL; }
    *n_out = 0;
    if (stats) *stats = fmb_stats{};
    FMB_TRY(check_scheme_args(n_searches, pi, l, u, partition));
    if (nq == 0) return FMB_OK;
    if (!packed) { set_error("packed is NULL"); return FMB_EINVAL; }
    std::vector<Job> jobs(1);
    Job& j = jobs[0];
    j.ix = ix; j.packed = packed; j.exc_pos = exc_pos; j.exc_sym = exc_sym; j.n_exc = n_exc; j.offsets = offsets; j.q_begin = 0; j.q_end = nq;
    j.edit = edit; j.n_searches = n_searches; j.n_parts = n_parts; j.pi = pi; j.l = l; j.u = u; j.partition = partition;
    j.out = out; j.capacity = capacity;
    return run_jobs(jobs, n_out, stats);
}

int fmb_search_and_locate_multi(const fmb_index* const* replicas, uint32_t n_replicas, const uint8_t* symbols, const uint64_t* offsets, uint64_t nq,
                                int edit, uint32_t n_searches, uint32_t n_parts, const uint32_t* pi, const uint32_t* l, const uint32_t* u,
                                const uint32_t* partition, fmb_loc32* out, uint64_t shard_capacity, uint64_t* n_out, fmb_stats* stats) {
    if (!replicas || n_replicas == 0 || !offsets || !n_out || (shard_capacity && !out)) { set_error("NULL argument"); return FMB_EINVAL; }
    for (uint32_t g = 0; g < n_replicas; ++g) {
        n_out[g] = 0;
        if (!replicas[g]) { set_error("replica %u is NULL", g); return FMB_EINVAL; }
        for (uint32_t h = 0; h < g; ++h)
            if (replicas[h] == replicas[g]) { set_error("replica %u listed twice", g); return FMB_EINVAL; }
    }
    if (stats) *stats = fmb_stats{};
    FMB_TRY(check_scheme_args(n_searches, pi, l, u, partition));
    if (nq == 0) return FMB_OK;
    // contiguous shards of ceil(nq / G) queries
    const uint64_t per = (nq + n_replicas - 1) / n_replicas;
    std::vector<Job> jobs(n_replicas);
    for (uint32_t g = 0; g < n_replicas; ++g) {
        Job& j = jobs[g];
        j.ix = replicas[g]; j.symbols = symbols; j.offsets = offsets;
        j.q_begin = std::min<uint64_t>(nq, (uint64_t)g * per); j.q_end = std::min<uint64_t>(nq, j.q_begin + per);
        j.edit = edit; j.n_searches = n_searches; j.n_parts = n_parts; j.pi = pi; j.l = l; j.u = u; j.partition = partition;
        j.out = out + (uint64_t)g * shard_capacity; j.capacity = shard_capacity;
    }
    return run_jobs(jobs, n_out, stats);
}

int fmb_search_and_locate_parts(const fmb_index* const* parts, uint32_t n_index_parts, const uint64_t* seq_base, const uint8_t* symbols,
                                const uint64_t* offsets, uint64_t nq, int edit, uint32_t n_searches, uint32_t n_parts, const uint32_t* pi,
                                const uint32_t* l, const uint32_t* u, const uint32_t* partition, fmb_loc32* out, uint64_t part_capacity,
                                uint64_t* n_out, fmb_stats* stats) {
    if (!parts || n_index_parts == 0 || !seq_base || !offsets || !n_out || (part_capacity && !out)) { set_error("NULL argument"); return FMB_EINVAL; }
    for (uint32_t g = 0; g < n_index_parts; ++g) {
        n_out[g] = 0;
        if (!parts[g]) { set_error("part %u is NULL", g); return FMB_EINVAL; }
        for (uint32_t h = 0; h < g; ++h)
            if (parts[h] == parts[g]) { set_error("part %u listed twice", g); return FMB_EINVAL; }
        if (seq_base[g] + parts[g]->n_delims > 0xFFFFFFFFull) { set_error("part %u: sequence numbers do not fit 32 bits", g); return FMB_EUNSUPPORTED; }
    }
    if (stats) *stats = fmb_stats{};
    FMB_TRY(check_scheme_args(n_searches, pi, l, u, partition));
    if (nq == 0) return FMB_OK;
    // every part searches the whole batch (the parts hold different sequences: an occurrence lies in exactly one of them)
    std::vector<Job> jobs(n_index_parts);
    for (uint32_t g = 0; g < n_index_parts; ++g) {
        Job& j = jobs[g];
        j.ix = parts[g]; j.symbols = symbols; j.offsets = offsets; j.q_begin = 0; j.q_end = nq;
        j.edit = edit; j.n_searches = n_searches; j.n_parts = n_parts; j.pi = pi; j.l = l; j.u = u; j.partition = partition;
        j.out = out + (uint64_t)g * part_capacity; j.capacity = part_capacity;
    }
    const int rc = run_jobs(jobs, n_out, stats);
    // sequence numbers of the whole collection
    for (uint32_t g = 0; g < n_index_parts; ++g) {
        const uint32_t base = (uint32_t)seq_base[g];
        if (base == 0) continue;
        fmb_loc32* rows = out + (uint64_t)g * part_capacity;
        const uint64_t n = std::min<uint64_t>(n_out[g], part_capacity);
        for (uint64_t i = 0; i < n; ++i) rows[i].seq += base;
    }
    return rc;
}

}  // extern "C"
