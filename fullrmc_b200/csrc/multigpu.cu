// multigpu.cu -- full_pairs_histograms_coords over several GPUs of one box, behind ONE call of the C ABI (SURVEY.md
// section 8b: frmc_full_pairs_hist(ndev, devs, ...) "shards + NCCL inside"; section 8e).
//
// An unmodified Engine is one Python process; its compute_data() must be able to use every GPU of the box without
// torchrun.  One call does:
//   1. device 0 orders the caller's atoms (devlayout.cu; raw arrays up once),
//   2. the 20 B/atom store is copied to the other devices over NVLink (cudaMemcpyPeerAsync, event-ordered),
//   3. every device lists and sweeps its share of the triangular row list (build_rows: rows dealt boustrophedon over
//      the shards), driven by one host thread per device (the list builder reads two sizes back),
//   4. ONE ncclAllReduce(sum) of the 64-bit ordered counts (2 * nEl^2 * histSize + 2 words: 400 KB at nEl = 5,
//      histSize = 1000) over NVLink / NVSwitch -- an in-process communicator made once per device set with
//      ncclCommInitAll, the library loaded at run time (libnccl.so.2: the process's own copy when torch brought one),
//   5. device 0 converts and returns the histograms.
// Integer counts: the result is identical for any number of devices (tests/test_multi_gpu.py).
#include "common.cuh"
#include "layout.h"

#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace frmc {

GridParams make_grid(float rmin, float rmax, float bin, int hs);
int full_hist_launch(cudaStream_t stream, int sm_count, int mode, const float4 *atoms, const uint32_t *orig,
                     int64_t npad, float4 *bbox, const WorkItem *rows, int n_rows, int n_pairs, PairLists &lists,
                     const int32_t *mol_by_orig, uint32_t mol_span,
                     const Lattice &L, const GridParams &g, int nEl, unsigned long long *counts, unsigned long long *stats);
void pack_rows(const std::vector<WorkItem> &rows, std::vector<unsigned char> &blob, int &n_pairs);
int launch_counts64_to_float(cudaStream_t stream, const unsigned long long *counts, float *out, long long cells2);
PairLists &stateless_lists_for(int dev);
extern int g_device_layout;

// ---------------------------------------------------------------- NCCL, loaded at run time
struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    bool ok = false;
};

static NcclApi &nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("FULLRMC_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (!nm || !*nm) continue;
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
#define FRMC_NCCL_SYM(field, sym) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym))
        FRMC_NCCL_SYM(CommInitAll, "ncclCommInitAll");
        FRMC_NCCL_SYM(CommDestroy, "ncclCommDestroy");
        FRMC_NCCL_SYM(AllReduce, "ncclAllReduce");
        FRMC_NCCL_SYM(GroupStart, "ncclGroupStart");
        FRMC_NCCL_SYM(GroupEnd, "ncclGroupEnd");
        FRMC_NCCL_SYM(GetErrorString, "ncclGetErrorString");
        FRMC_NCCL_SYM(GetVersion, "ncclGetVersion");
#undef FRMC_NCCL_SYM
        api.ok = api.CommInitAll && api.AllReduce && api.GroupStart && api.GroupEnd && api.GetErrorString;
    });
    return api;
}

struct CommSet {
    std::vector<int> devs;
    std::vector<ncclComm_t> comms;
};

// one communicator set per distinct device list, made once per process
static int get_comms(const std::vector<int> &devs, CommSet **out)
{
    static std::mutex mu;
    static std::vector<CommSet *> sets;
    std::lock_guard<std::mutex> lock(mu);
    for (CommSet *s : sets)
        if (s->devs == devs) { *out = s; return FRMC_OK; }
    NcclApi &api = nccl_api();
    FRMC_REQUIRE(api.ok, FRMC_ECUDA, "libnccl.so.2 could not be loaded (%s); set FULLRMC_B200_NCCL to its path", dlerror() ? dlerror() : "symbols missing");
    CommSet *s = new CommSet;
    s->devs = devs;
    s->comms.resize(devs.size());
    ncclResult_t r = api.CommInitAll(s->comms.data(), (int)devs.size(), devs.data());
    if (r != ncclSuccess) {
        set_error("ncclCommInitAll over %zu devices failed: %s", devs.size(), api.GetErrorString(r));
        delete s;
        return FRMC_ECUDA;
    }
    sets.push_back(s);
    *out = s;
    return FRMC_OK;
}

static char g_reduce_path[64] = "none";

}  // namespace frmc

using namespace frmc;

extern "C" const char *frmc_multi_reduce_path(void) { return g_reduce_path; }

extern "C" int frmc_full_pairs_histograms_coords_multi(int ndev, const int *devs, const float *coords, int64_t n, const float *basis,
                                                       int isPBC, const int32_t *mol, const int32_t *el, int nEl, float rmin,
                                                       float rmax, float bin, int hs, float *hintra, float *hinter,
                                                       uint64_t *edge_overflow)
{
    FRMC_REQUIRE(ndev >= 1 && ndev <= 64 && devs, FRMC_EINVAL, "bad device list (%d devices)", ndev);
    FRMC_REQUIRE(n >= 0, FRMC_EINVAL, "negative atom count");
    FRMC_REQUIRE(n == 0 || (coords && mol && el), FRMC_EINVAL, "NULL input array");
    FRMC_REQUIRE(hs >= 1 && hintra && hinter, FRMC_EINVAL, "bad histogram arguments");
    FRMC_REQUIRE(nEl >= 1 && nEl <= FRMC_MAX_ELEMENTS, FRMC_ELIMIT, "numberOfElements %d outside 1..%d", nEl, FRMC_MAX_ELEMENTS);
    for (int a = 0; a < ndev; ++a)
        for (int b = a + 1; b < ndev; ++b) FRMC_REQUIRE(devs[a] != devs[b], FRMC_EINVAL, "device %d listed twice", devs[a]);
    const int64_t cells = (int64_t)nEl * nEl * hs;
    std::vector<DeviceCtx *> ctx((size_t)ndev);
    for (int k = 0; k < ndev; ++k) {
        ctx[(size_t)k] = get_ctx(devs[k]);
        if (!ctx[(size_t)k]) return FRMC_ECUDA;
    }
    DeviceCtx *c0 = ctx[0];
    FRMC_CUDA(cudaSetDevice(c0->dev));
    const bool timing = getenv("FRMC_MULTI_TIMING") != nullptr;     // debug: host wall clock at the phase boundaries (stderr)
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto t_start = now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        cudaStreamSynchronize(c0->stream);
        fprintf(stderr, "[multi x%d] %-28s %8.3f ms\n", ndev, what, std::chrono::duration<double, std::milli>(now() - t_start).count());
    };

    // 1. the store layout on device 0 (the scratch is per calling thread; the worker threads below must see THIS
    //    thread's instance, hence the reference)
    static thread_local HostLayout lay_tls;
    HostLayout &lay = lay_tls;
    float4 *d_atoms0 = nullptr;
    uint32_t *d_orig0 = nullptr;
    int32_t *d_keys0 = nullptr;
    int rc;
    // direct NVLink copies device 0 -> device k (without peer access cudaMemcpyPeerAsync stages through the host)
    {
        static std::mutex mu;
        static bool enabled[64][64];
        std::lock_guard<std::mutex> lock(mu);
        for (int k = 1; k < ndev; ++k) {
            const int a = c0->dev, b = ctx[(size_t)k]->dev;
            if (a >= 64 || b >= 64 || enabled[a][b]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, a, b) == cudaSuccess && can) {
                cudaSetDevice(a);
                if (cudaDeviceEnablePeerAccess(b, 0) != cudaSuccess) cudaGetLastError();      // already enabled by the host program
                cudaSetDevice(b);
                if (cudaDeviceEnablePeerAccess(a, 0) != cudaSuccess) cudaGetLastError();
            }
            enabled[a][b] = true;
        }
    }
    FRMC_CUDA(cudaSetDevice(c0->dev));
    if (g_device_layout) {
        // The raw arrays (20 B/atom) reach device 0 over EVERY device's own PCIe link: device k stages slice k in its
        // page-locked scratch, copies it up on its stream and forwards it to device 0 over NVLink; device 0's stream waits
        // for the slices and orders the atoms (one link: 1.0 ms of the call at 10^6 atoms, the largest serial piece).
        bool raw_on_device = false;
        static const bool sliced_on = []() { const char *e = getenv("FRMC_SLICED_UPLOAD"); return !(e && atoi(e) == 0); }();
        // (measured at 10^6 atoms: 4.1 -> 3.7 ms per call on 8 GPUs, no gain on 4, 0.5 ms SLOWER on 2 -- one staging thread per
        //  device and the extra NVLink hop against four pipelined staging threads on one link -- hence from 5 devices on)
        if (sliced_on && ndev >= 5 && n >= 65536 * (int64_t)ndev) {
            float *d_c0 = (float *)ctx_buffer(c0, 7, sizeof(float) * 3 * (size_t)n);
            int32_t *d_e0 = (int32_t *)ctx_buffer(c0, 8, sizeof(int32_t) * 2 * (size_t)n);
            if (!d_c0 || !d_e0) return FRMC_ENOMEM;
            int32_t *d_k0 = d_e0 + n;
            std::vector<cudaEvent_t> up((size_t)ndev, nullptr);
            std::vector<int> errs((size_t)ndev, 0);
            std::vector<std::thread> th;
            for (int k = 0; k < ndev; ++k)
                th.emplace_back([&, k] {
                    DeviceCtx *c = ctx[(size_t)k];
                    if (cudaSetDevice(c->dev) != cudaSuccess) { errs[(size_t)k] = 1; return; }
                    const int64_t a = n * k / ndev, b = n * (k + 1) / ndev;
                    const size_t cnt = (size_t)(b - a), bc = sizeof(float) * 3 * cnt, bi = sizeof(int32_t) * cnt;
                    unsigned char *pin = (unsigned char *)ctx_pinned(c, bc + 2 * bi);
                    unsigned char *tmp = (k == 0) ? nullptr : (unsigned char *)ctx_buffer(c, 7, bc + 2 * bi);
                    if (!pin || (k > 0 && !tmp)) { errs[(size_t)k] = 1; return; }
                    memcpy(pin, coords + 3 * a, bc);
                    memcpy(pin + bc, el + a, bi);
                    memcpy(pin + bc + bi, mol + a, bi);
                    bool ok = true;
                    if (k == 0) {
                        ok = ok && cudaMemcpyAsync(d_c0 + 3 * a, pin, bc, cudaMemcpyHostToDevice, c->stream) == cudaSuccess;
                        ok = ok && cudaMemcpyAsync(d_e0 + a, pin + bc, bi, cudaMemcpyHostToDevice, c->stream) == cudaSuccess;
                        ok = ok && cudaMemcpyAsync(d_k0 + a, pin + bc + bi, bi, cudaMemcpyHostToDevice, c->stream) == cudaSuccess;
                    } else {
                        ok = ok && cudaMemcpyAsync(tmp, pin, bc + 2 * bi, cudaMemcpyHostToDevice, c->stream) == cudaSuccess;
                        ok = ok && cudaMemcpyPeerAsync(d_c0 + 3 * a, c0->dev, tmp, c->dev, bc, c->stream) == cudaSuccess;
                        ok = ok && cudaMemcpyPeerAsync(d_e0 + a, c0->dev, tmp + bc, c->dev, bi, c->stream) == cudaSuccess;
                        ok = ok && cudaMemcpyPeerAsync(d_k0 + a, c0->dev, tmp + bc + bi, c->dev, bi, c->stream) == cudaSuccess;
                        ok = ok && cudaEventCreateWithFlags(&up[(size_t)k], cudaEventDisableTiming) == cudaSuccess;
                        ok = ok && cudaEventRecord(up[(size_t)k], c->stream) == cudaSuccess;
                    }
                    if (!ok) errs[(size_t)k] = 1;
                });
            for (auto &t : th) t.join();
            FRMC_CUDA(cudaSetDevice(c0->dev));
            bool bad = false;
            for (int k = 0; k < ndev; ++k) bad = bad || errs[(size_t)k];
            for (int k = 1; k < ndev && !bad; ++k) bad = bad || cudaStreamWaitEvent(c0->stream, up[(size_t)k], 0) != cudaSuccess;
            for (int k = 1; k < ndev; ++k) if (up[(size_t)k]) cudaEventDestroy(up[(size_t)k]);
            FRMC_REQUIRE(!bad, FRMC_ECUDA, "sliced upload of the atom arrays failed: %s", cudaGetErrorString(cudaGetLastError()));
            raw_on_device = true;
        }
        rc = device_layout(c0, coords, n, mol, el, nEl, isPBC, lay, &d_atoms0, &d_orig0, &d_keys0, raw_on_device);
        if (rc) return rc;
    } else {
        rc = build_layout(coords, n, mol, el, nEl, isPBC, lay);
        if (rc) return rc;
        d_atoms0 = (float4 *)ctx_buffer(c0, 0, sizeof(float4) * (size_t)std::max<int64_t>(lay.npad, 1));
        d_orig0 = (uint32_t *)ctx_buffer(c0, 1, sizeof(uint32_t) * (size_t)std::max<int64_t>(lay.npad, 1));
        if (!d_atoms0 || !d_orig0) return FRMC_ENOMEM;
        if (lay.npad > 0) {
            FRMC_CUDA(cudaMemcpyAsync(d_atoms0, lay.rec.data(), sizeof(float4) * (size_t)lay.npad, cudaMemcpyHostToDevice, c0->stream));
            FRMC_CUDA(cudaMemcpyAsync(d_orig0, lay.orig.data(), sizeof(uint32_t) * (size_t)lay.npad, cudaMemcpyHostToDevice, c0->stream));
        }
    }
    lap("layout on device 0");
    Lattice L;
    for (int i = 0; i < 9; ++i) L.b[i] = basis ? basis[i] : ((i % 4 == 0) ? 1.0f : 0.0f);
    const GridParams g = make_grid(rmin, rmax, bin, hs);
    const int mode = choose_mode_from_bounds(L.b, isPBC, lay.lo, lay.hi);

    // 2. per-device buffers; the store travels device 0 -> device k behind device 0's stream
    struct Dev {
        float4 *atoms; uint32_t *orig; WorkItem *rows; unsigned long long *counts; float4 *bbox; int32_t *mol;
        std::vector<unsigned char> blob; int n_rows, n_pairs; int rc; std::string err;
    };
    std::vector<Dev> D((size_t)ndev);
    const size_t n_words = (size_t)(2 * cells + 3);
    for (int k = 0; k < ndev; ++k) {
        Dev &d = D[(size_t)k];
        DeviceCtx *c = ctx[(size_t)k];
        FRMC_CUDA(cudaSetDevice(c->dev));
        std::vector<WorkItem> rows;
        build_rows(lay, 1, k, ndev, rows);
        pack_rows(rows, d.blob, d.n_pairs);
        d.n_rows = (int)rows.size();
        d.atoms = (k == 0) ? d_atoms0 : (float4 *)ctx_buffer(c, 0, sizeof(float4) * (size_t)std::max<int64_t>(lay.npad, 1));
        d.orig = (k == 0) ? d_orig0 : (uint32_t *)ctx_buffer(c, 1, sizeof(uint32_t) * (size_t)std::max<int64_t>(lay.npad, 1));
        d.rows = (WorkItem *)ctx_buffer(c, 2, d.blob.size());
        d.bbox = (float4 *)ctx_buffer(c, 3, sizeof(float4) * 18 * (size_t)(lay.npad / SEG_PAD + 1));
        d.counts = (unsigned long long *)ctx_buffer(c, 4, sizeof(unsigned long long) * n_words);
        d.mol = nullptr;
        if (lay.mol_span > 0) d.mol = (int32_t *)ctx_buffer(c, 6, sizeof(int32_t) * (size_t)std::max<int64_t>(n, 1));
        if (!d.atoms || !d.orig || !d.rows || !d.bbox || !d.counts || (lay.mol_span > 0 && !d.mol)) return FRMC_ENOMEM;
        d.rc = FRMC_OK;
    }
    FRMC_CUDA(cudaSetDevice(c0->dev));
    cudaEvent_t ready;
    FRMC_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    if (lay.npad > 0)
        for (int k = 1; k < ndev; ++k) {
            FRMC_CUDA(cudaMemcpyPeerAsync(D[(size_t)k].atoms, ctx[(size_t)k]->dev, d_atoms0, c0->dev, sizeof(float4) * (size_t)lay.npad, c0->stream));
            FRMC_CUDA(cudaMemcpyPeerAsync(D[(size_t)k].orig, ctx[(size_t)k]->dev, d_orig0, c0->dev, sizeof(uint32_t) * (size_t)lay.npad, c0->stream));
        }
    FRMC_CUDA(cudaEventRecord(ready, c0->stream));
    lap("rows + peer copies");

    // 3. every device lists and sweeps its rows (one host thread each: the list builder synchronises its stream)
    auto work = [&](int k) {
        Dev &d = D[(size_t)k];
        DeviceCtx *c = ctx[(size_t)k];
        auto fail = [&](int code) { d.rc = code; d.err = last_error(); };
        if (cudaSetDevice(c->dev) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", c->dev); return fail(FRMC_ECUDA); }
        if (k > 0 && cudaStreamWaitEvent(c->stream, ready, 0) != cudaSuccess) { set_error("cudaStreamWaitEvent failed"); return fail(FRMC_ECUDA); }
        if (cudaMemsetAsync(d.counts, 0, sizeof(unsigned long long) * n_words, c->stream) != cudaSuccess) { set_error("memset failed"); return fail(FRMC_ECUDA); }
        if (d.mol && cudaMemcpyAsync(d.mol, mol, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) {
            set_error("molecule index upload failed"); return fail(FRMC_ECUDA);
        }
        if (d.n_rows > 0) {
            if (cudaMemcpyAsync(d.rows, d.blob.data(), d.blob.size(), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) {
                set_error("row list upload failed"); return fail(FRMC_ECUDA);
            }
            const int r = full_hist_launch(c->stream, c->sm_count, mode, d.atoms, d.orig, lay.npad, d.bbox, d.rows, d.n_rows, d.n_pairs,
                                           stateless_lists_for(c->dev), d.mol, lay.mol_span, L, g, nEl, d.counts, d.counts + 2 * cells);
            if (r) return fail(r);
        }
    };
    if (ndev == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int k = 0; k < ndev; ++k) th.emplace_back(work, k);
        for (auto &t : th) t.join();
    }
    for (int k = 0; k < ndev; ++k)
        if (D[(size_t)k].rc) { set_error("device %d: %s", ctx[(size_t)k]->dev, D[(size_t)k].err.c_str()); cudaEventDestroy(ready); return D[(size_t)k].rc; }

    lap("sweeps issued (threads joined)");
    // 4. one all-reduce of the 64-bit counts (+ the overflow and swept counters behind them)
    if (ndev > 1) {
        CommSet *cs = nullptr;
        std::vector<int> dl(devs, devs + ndev);
        rc = get_comms(dl, &cs);
        if (rc) { cudaEventDestroy(ready); return rc; }
        NcclApi &api = nccl_api();
        ncclResult_t r = api.GroupStart();
        for (int k = 0; k < ndev && r == ncclSuccess; ++k)
            r = api.AllReduce(D[(size_t)k].counts, D[(size_t)k].counts, 2 * (size_t)cells + 2, ncclUint64, ncclSum, cs->comms[(size_t)k], ctx[(size_t)k]->stream);
        const ncclResult_t r2 = api.GroupEnd();
        if (r != ncclSuccess || r2 != ncclSuccess) {
            set_error("ncclAllReduce failed: %s", api.GetErrorString(r != ncclSuccess ? r : r2));
            cudaEventDestroy(ready);
            return FRMC_ECUDA;
        }
        int ver = 0;
        if (api.GetVersion) api.GetVersion(&ver);
        snprintf(g_reduce_path, sizeof(g_reduce_path), "nccl %d x%d", ver, ndev);
    } else {
        snprintf(g_reduce_path, sizeof(g_reduce_path), "single device");
    }

    lap("all-reduce issued");
    // 5. device 0 returns the result; the other devices only have to finish
    FRMC_CUDA(cudaSetDevice(c0->dev));
    float *d_out = (float *)ctx_buffer(c0, 5, sizeof(float) * 2 * (size_t)cells);
    if (!d_out) { cudaEventDestroy(ready); return FRMC_ENOMEM; }
    rc = launch_counts64_to_float(c0->stream, D[0].counts, d_out, 2 * cells);
    if (rc) { cudaEventDestroy(ready); return rc; }
    unsigned long long ov = 0;
    FRMC_CUDA(cudaMemcpyAsync(hintra, d_out, sizeof(float) * (size_t)cells, cudaMemcpyDeviceToHost, c0->stream));
    FRMC_CUDA(cudaMemcpyAsync(hinter, d_out + cells, sizeof(float) * (size_t)cells, cudaMemcpyDeviceToHost, c0->stream));
    FRMC_CUDA(cudaMemcpyAsync(&ov, D[0].counts + 2 * cells, sizeof(ov), cudaMemcpyDeviceToHost, c0->stream));
    for (int k = 0; k < ndev; ++k) {
        FRMC_CUDA(cudaSetDevice(ctx[(size_t)k]->dev));
        FRMC_CUDA(cudaStreamSynchronize(ctx[(size_t)k]->stream));
    }
    FRMC_CUDA(cudaSetDevice(c0->dev));
    lap("all devices done, result home");
    cudaEventDestroy(ready);
    if (edge_overflow) *edge_overflow = ov;
    return FRMC_OK;
}
