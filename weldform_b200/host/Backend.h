// One or several GPUs behind the host mirror of SPH::Domain (weldform_b200/host/Domain.h).
//
// WELDFORM_GPUS=N (default 1) runs a driver on N GPUs of one node from ONE process: N wf_ctx, one per device, each driven by its own host
// thread (the per-phase entry points of the C ABI are collective on several GPUs: every rank must be inside the same call at the same time),
// the library's own slab planner (wf_plan_slabs_ex) and NCCL communicator (wf_comm_unique_id / wf_comm_init). Domain::Particles stays GLOBAL
// and in creation order: this class splits the uploads by owner rank and scatters the downloads back through each rank's global-index list,
// which it refreshes whenever particles migrated between ranks at a neighbour rebuild (wf_comm_info).
// The reference is a single-process OpenMP code (no counterpart); with one GPU everything here degenerates to direct calls.
#ifndef WF_HOST_BACKEND_H
#define WF_HOST_BACKEND_H

#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "weldform_b200.h"
#include "fatal.h"

namespace wf_host {

// persistent workers: rank 0 runs on the calling thread, ranks 1..n-1 on their own threads for the lifetime of a Solve call
class RankPool {
public:
    ~RankPool() { shutdown(); }
    void start(int nranks)
    {
        shutdown(); n = nranks;
        for (int r = 1; r < n; ++r) th.emplace_back([this, r] { worker(r); });
    }
    void run(const std::function<void(int)> & f)
    {
        if (n <= 1) { f(0); return; }
        { std::lock_guard<std::mutex> lk(m); job = f; pending = n - 1; ++gen; }
        cv_go.notify_all();
        f(0);
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [this] { return pending == 0; });
    }
    void shutdown()
    {
        { std::lock_guard<std::mutex> lk(m); stop = true; }
        cv_go.notify_all();
        for (auto & t : th) t.join();
        th.clear(); stop = false; n = 1;
    }
private:
    void worker(int r)
    {
        unsigned long seen = 0;
        for (;;) {
            std::function<void(int)> j;
            { std::unique_lock<std::mutex> lk(m); cv_go.wait(lk, [&] { return stop || gen != seen; }); if (stop) return; seen = gen; j = job; }
            j(r);
            { std::lock_guard<std::mutex> lk(m); if (--pending == 0) cv_done.notify_one(); }
        }
    }
    std::vector<std::thread> th; std::mutex m; std::condition_variable cv_go, cv_done;
    std::function<void(int)> job; unsigned long gen = 0; int pending = 0, n = 1; bool stop = false;
};

class Backend {
public:
    int n = 0;                                       // GPUs (ranks); 0 = not created
    std::vector<wf_ctx *> ctx;
    int64_t n_total = 0, n_sph = 0;                  // particles of the whole domain (SPH + rigid-mesh element particles), SPH particles
    int steps_hint = 0; double time_hint = 0.0;      // for error messages

    ~Backend() { destroy(); }
    bool active() const { return n > 0; }
    wf_ctx * c0() const { return ctx.empty() ? nullptr : ctx[0]; }

    // collective call: f(rank) on every rank at the same time; the first failing rank's message becomes the Fatal
    void each(const char * what, const std::function<int(int)> & f)
    {
        std::vector<int> rc((size_t)n, 0);
        pool.run([&](int r) { rc[(size_t)r] = f(r); });
        for (int r = 0; r < n; ++r)
            if (rc[(size_t)r] != 0) throw new Fatal("%s failed (%d) on GPU rank %d at step %d, time %g: %s", what, rc[(size_t)r], r, steps_hint, time_hint, wf_last_error(ctx[(size_t)r]));
    }
    void phase(const char * what, int (*fn)(wf_ctx *)) { each(what, [&](int r) { return fn(ctx[(size_t)r]); }); after_collective(); }
    void phase_i(const char * what, int (*fn)(wf_ctx *, int), int arg) { each(what, [&](int r) { return fn(ctx[(size_t)r], arg); }); after_collective(); }

    void create(int device0, int ngpus)
    {
        destroy();
        if (ngpus < 1) ngpus = 1;
        int64_t sz[5] = {0, 0, 0, 0, 0}; wf_abi_sizes(sz);
        if (sz[0] != (int64_t)sizeof(wf_params) || sz[1] != (int64_t)sizeof(wf_material) || sz[2] != (int64_t)sizeof(wf_particles) || sz[3] != (int64_t)sizeof(wf_mesh) || sz[4] != (int64_t)sizeof(wf_scalars))
            throw new Fatal("weldform_b200: libweldform_b200.so was built from another include/weldform_b200.h than this program (structure sizes differ): rebuild both");
        ctx.assign((size_t)ngpus, nullptr);
        for (int r = 0; r < ngpus; ++r)
            if (wf_create(device0 + r, &ctx[(size_t)r]) != 0) { std::string e = wf_last_error(nullptr); destroy_ctxs(); throw new Fatal("weldform_b200: GPU %d: %s", device0 + r, e.c_str()); }
        n = ngpus;
        pool.start(n);
        gidx.assign((size_t)n, std::vector<int64_t>()); mig_seen.assign((size_t)n, -1);
        if (n > 1) {
            unsigned char id[128];
            if (wf_comm_unique_id(id) != 0) throw new Fatal("weldform_b200: wf_comm_unique_id failed (NCCL not available?)");
            each("wf_comm_init", [&](int r) { return wf_comm_init(ctx[(size_t)r], r, n, id); });
        }
    }
    void destroy()
    {
        if (n > 1) pool.run([&](int r) { wf_destroy(ctx[(size_t)r]); ctx[(size_t)r] = nullptr; });   // NCCL teardown from every rank's own thread
        pool.shutdown();
        destroy_ctxs();
        n = 0; gidx.clear(); mig_seen.clear(); groups.clear();
    }

    // ---- set-up -----------------------------------------------------------------------------------------------------
    // g: the whole domain in creation order. plan_x / plan_h: positions and smoothing lengths of ALL particles (SPH first, then the element
    // particles of the rigid meshes), which shape the cell grid the slabs are cut from (Neighbour.cpp:10-139)
    void set_particles(const wf_particles & g, const double * plan_x, const double * plan_h, int64_t n_all, int kernel_type, const double dom_min[3], const double dom_max[3])
    {
        n_sph = g.n_sph; n_total = n_all;
        if (n == 1) {
            each("wf_set_particles", [&](int) { return wf_set_particles(ctx[0], &g); });
            gidx[0].resize((size_t)n_all); for (int64_t i = 0; i < n_all; ++i) gidx[0][(size_t)i] = i;
            return;
        }
        std::vector<int32_t> rank_of((size_t)n_all), bounds((size_t)n + 1);
        if (wf_plan_slabs_ex(plan_x, plan_h, n_all, g.n_sph, kernel_type, dom_min, dom_max, n, rank_of.data(), bounds.data()) != 0)
            throw new Fatal("weldform_b200: wf_plan_slabs_ex failed (fewer cell layers than GPUs?)");
        std::vector<std::vector<int64_t> > sel((size_t)n);
        for (int64_t i = 0; i < g.n_sph; ++i) { const int r = rank_of[(size_t)i]; if (r < 0 || r >= n) throw new Fatal("weldform_b200: slab planner left particle %lld unassigned", (long long)i); sel[(size_t)r].push_back(i); }
        each("wf_set_particles", [&](int r) {
            const std::vector<int64_t> & s = sel[(size_t)r]; const size_t m = s.size();
            std::vector<std::vector<double> > dbuf; std::vector<std::vector<int32_t> > ibuf;
            dbuf.reserve(32); ibuf.reserve(16);
            wf_particles p; std::memset(&p, 0, sizeof(p));
            auto cut = [&](const double * src, int nc) -> const double * {
                if (!src) return nullptr;
                dbuf.emplace_back(m * (size_t)nc); std::vector<double> & d = dbuf.back();
                for (size_t k = 0; k < m; ++k) for (int c = 0; c < nc; ++c) d[k * nc + c] = src[(size_t)s[k] * nc + c];
                return d.data();
            };
            auto cuti = [&](const int32_t * src) -> const int32_t * {
                if (!src) return nullptr;
                ibuf.emplace_back(m); std::vector<int32_t> & d = ibuf.back();
                for (size_t k = 0; k < m; ++k) d[k] = src[(size_t)s[k]];
                return d.data();
            };
            p.n_sph = (int64_t)m; p.x = cut(g.x, 3); p.v = cut(g.v, 3); p.mass = cut(g.mass, 1); p.density = cut(g.density, 1); p.h = cut(g.h, 1); p.ref_density = cut(g.ref_density, 1);
            p.id = cuti(g.id); p.is_free = cuti(g.is_free); p.not_write_surf_id = cuti(g.not_write_surf_id);
            p.G = cut(g.G, 1); p.Cs = cut(g.Cs, 1); p.P0 = cut(g.P0, 1); p.alpha = cut(g.alpha, 1); p.beta = cut(g.beta, 1); p.TI = cut(g.TI, 1); p.TIn = cut(g.TIn, 1);
            p.TIInitDist = cut(g.TIInitDist, 1); p.Ep = cut(g.Ep, 1); p.k_T = cut(g.k_T, 1); p.cp_T = cut(g.cp_T, 1); p.h_conv = cut(g.h_conv, 1); p.T_inf = cut(g.T_inf, 1);
            p.th_exp = cut(g.th_exp, 1); p.q_source = cut(g.q_source, 1);
            p.pres_eq = cuti(g.pres_eq); p.fail = cuti(g.fail); p.material = cuti(g.material); p.material_model = cuti(g.material_model); p.mat_id = cuti(g.mat_id); p.thermal_bc = cuti(g.thermal_bc);
            p.sigmay = cut(g.sigmay, 1); p.T = cut(g.T, 1);
            p.global_index = s.data(); p.n_sph_global = g.n_sph;
            return wf_set_particles(ctx[(size_t)r], &p);
        });
        for (int r = 0; r < n; ++r) mig_seen[(size_t)r] = -1;       // ownership lists come from the device once the solver is initialised
    }

    // ---- ownership: creation index of every local particle of every rank (ascending), refreshed after migrations -------------------------
    void after_collective()
    {
        if (n <= 1) return;
        bool changed = false;
        for (int r = 0; r < n; ++r) {
            int64_t info[8]; if (wf_comm_info(ctx[(size_t)r], info) != 0) continue;
            const int64_t sig = info[2] * 1000003 + info[3];
            if (sig != mig_seen[(size_t)r]) changed = true;
        }
        if (changed) refresh_ownership();
    }
    void refresh_ownership()
    {
        if (n <= 1) return;
        each("wf_download(global_index)", [&](int r) {
            const int64_t nl = wf_num_particles(ctx[(size_t)r]);
            std::vector<int32_t> gi((size_t)nl);
            const int rc = nl ? wf_download(ctx[(size_t)r], WF_F_GLOBAL_INDEX, gi.data()) : 0;
            gidx[(size_t)r].assign(gi.begin(), gi.end());
            int64_t info[8]; if (wf_comm_info(ctx[(size_t)r], info) == 0) mig_seen[(size_t)r] = info[2] * 1000003 + info[3];
            return rc;
        });
        for (auto & g : groups) register_group(g);                   // local indices went stale
        ++ownership_epoch;
    }
    unsigned long ownership_epoch = 0;

    // ---- data movement in GLOBAL creation order ----------------------------------------------------------------------------------------
    // whole domain: out[i*nc + c] for every particle i (rigid-mesh element particles are replicated on every rank: rank 0's copy is taken)
    void download_all(wf_field f, int nc, bool is_int, void * out)
    {
        if (n == 1) { each("wf_download", [&](int) { return wf_download(ctx[0], f, out); }); return; }
        const size_t es = is_int ? 4 : 8;
        each("wf_download", [&](int r) {
            const std::vector<int64_t> & gi = gidx[(size_t)r]; if (gi.empty()) return 0;
            std::vector<char> tmp(gi.size() * (size_t)nc * es);
            const int rc = wf_download(ctx[(size_t)r], f, tmp.data()); if (rc != 0) return rc;
            for (size_t k = 0; k < gi.size(); ++k) { if (gi[k] >= n_sph && r != 0) continue; std::memcpy((char *)out + (size_t)gi[k] * nc * es, tmp.data() + k * nc * es, (size_t)nc * es); }
            return 0;
        });
    }
    void upload_all(wf_field f, int nc, bool is_int, const void * in)
    {
        if (n == 1) { each("wf_upload", [&](int) { return wf_upload(ctx[0], f, in); }); return; }
        const size_t es = is_int ? 4 : 8;
        each("wf_upload", [&](int r) {
            const std::vector<int64_t> & gi = gidx[(size_t)r]; if (gi.empty()) return 0;
            std::vector<char> tmp(gi.size() * (size_t)nc * es);
            for (size_t k = 0; k < gi.size(); ++k) std::memcpy(tmp.data() + k * nc * es, (const char *)in + (size_t)gi[k] * nc * es, (size_t)nc * es);
            return wf_upload(ctx[(size_t)r], f, tmp.data());
        });
    }

    // index groups (creation indices of SPH particles, e.g. the rows a callback touches): registered on the owner ranks' devices
    struct Group { std::vector<int32_t> idx; std::vector<std::vector<int32_t> > pos; std::vector<int> handle; };    // pos[r][k]: position in idx of the k-th local member of rank r
    int add_group(const std::vector<int32_t> & idx)
    {
        groups.emplace_back(); Group & g = groups.back(); g.idx = idx; g.pos.assign((size_t)n, std::vector<int32_t>()); g.handle.assign((size_t)n, -1);
        register_group(g);
        return (int)groups.size() - 1;
    }
    size_t group_size(int gi) const { return groups[(size_t)gi].idx.size(); }
    void download_group(int gi, wf_field f, int nc, bool is_int, void * out)
    {
        Group & g = groups[(size_t)gi]; const size_t es = is_int ? 4 : 8;
        each("wf_download_registered", [&](int r) {
            const std::vector<int32_t> & pos = g.pos[(size_t)r]; if (pos.empty()) return 0;
            if (n == 1) return wf_download_registered(ctx[0], f, g.handle[0], out);
            std::vector<char> tmp(pos.size() * (size_t)nc * es);
            const int rc = wf_download_registered(ctx[(size_t)r], f, g.handle[(size_t)r], tmp.data()); if (rc != 0) return rc;
            for (size_t k = 0; k < pos.size(); ++k) std::memcpy((char *)out + (size_t)pos[k] * nc * es, tmp.data() + k * nc * es, (size_t)nc * es);
            return 0;
        });
    }
    void upload_group(int gi, wf_field f, int nc, bool is_int, const void * in)
    {
        Group & g = groups[(size_t)gi]; const size_t es = is_int ? 4 : 8;
        each("wf_upload_registered", [&](int r) {
            const std::vector<int32_t> & pos = g.pos[(size_t)r]; if (pos.empty()) return 0;
            if (n == 1) return wf_upload_registered(ctx[0], f, g.handle[0], in);
            std::vector<char> tmp(pos.size() * (size_t)nc * es);
            for (size_t k = 0; k < pos.size(); ++k) std::memcpy(tmp.data() + k * nc * es, (const char *)in + (size_t)pos[k] * nc * es, (size_t)nc * es);
            return wf_upload_registered(ctx[(size_t)r], f, g.handle[(size_t)r], tmp.data());
        });
    }
    void fill_group(int gi, wf_field f, const double * value)
    {
        Group & g = groups[(size_t)gi];
        each("wf_fill_registered", [&](int r) { return g.handle[(size_t)r] < 0 ? 0 : wf_fill_registered(ctx[(size_t)r], f, g.handle[(size_t)r], value); });
    }

private:
    RankPool pool;
    std::vector<std::vector<int64_t> > gidx;         // per rank: creation index of local particle k
    std::vector<int64_t> mig_seen;
    std::vector<Group> groups;

    void destroy_ctxs() { for (auto & c : ctx) if (c) { wf_destroy(c); c = nullptr; } ctx.clear(); }
    void register_group(Group & g)
    {
        // owner rank and local index of every member: local index = position in the rank's ascending creation-index list
        for (int r = 0; r < n; ++r) {
            if (g.handle[(size_t)r] >= 0) { wf_subset_release(ctx[(size_t)r], g.handle[(size_t)r]); g.handle[(size_t)r] = -1; }
            g.pos[(size_t)r].clear();
        }
        std::vector<std::vector<int32_t> > local((size_t)n);
        if (n == 1) { local[0] = g.idx; g.pos[0].resize(g.idx.size()); for (size_t k = 0; k < g.idx.size(); ++k) g.pos[0][k] = (int32_t)k; }
        else {
            // gidx[r] is ascending: binary search per member
            for (size_t k = 0; k < g.idx.size(); ++k) {
                for (int r = 0; r < n; ++r) {
                    const std::vector<int64_t> & gi = gidx[(size_t)r];
                    size_t lo = 0, hi = gi.size();
                    while (lo < hi) { const size_t mid = (lo + hi) / 2; if (gi[mid] < g.idx[k]) lo = mid + 1; else hi = mid; }
                    if (lo < gi.size() && gi[lo] == g.idx[k]) { local[(size_t)r].push_back((int32_t)lo); g.pos[(size_t)r].push_back((int32_t)k); break; }
                }
            }
        }
        each("wf_subset_register", [&](int r) {
            if (local[(size_t)r].empty()) return 0;
            return wf_subset_register(ctx[(size_t)r], local[(size_t)r].data(), (int64_t)local[(size_t)r].size(), &g.handle[(size_t)r]);
        });
    }
};

} // namespace wf_host
#endif
