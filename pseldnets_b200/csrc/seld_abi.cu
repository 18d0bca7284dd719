// C-ABI layer of libseldfeat.so: plan construction (host-side table building, device upload) and
// the per-call argument checks + launches.  See include/seldfeat.h for the contract.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <new>
#include <vector>

#include "../../include/seldfeat.h"
#include "seld_plan.h"

struct HostPipe {                 // device slots + streams of seld_logmel_iv_f32_host
    static const int kSlots = 3;
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    cudaEvent_t ev_start = nullptr, ev_in[kSlots] = {}, ev_k[kSlots] = {}, ev_out[kSlots] = {};
    float* din[kSlots] = {};
    float* dout[kSlots] = {};
    size_t cap_in = 0, cap_out = 0;
    bool ready = false;
};

struct seld_plan {
    HostPipe pipe;
    seld::PlanDev dev;
    int device;
    int sm_count;
    int n_fft;
    size_t smem_optin;
    seld::MelTiles mt;     // mel bank as tensor-core operand (iv5)
    int iv_kernel;         // 5 / 3 / 2: generation of the 4-channel IV kernel in use; 0: general kernel only
    bool use_iv2;          // a fused 4-channel IV kernel (iv2 / iv3 / iv5) is in use
    bool iv2_ok;           // the fp32 mel-walk kernel can take this bank (fallback of iv5 for unaligned outputs)
    bool lm4_ok;           // log-mel-only mode of the iv2 kernel available (channels beyond the first four, Logmel_Extractor)
    void* blob;            // one device allocation holding every table
    int* redo_pool;        // kRedoPairs pairs of ints (inside the blob), zero between launches: see FoaArgs::redo_flags
    mutable unsigned redo_next;
};
constexpr unsigned kRedoPairs = 256;
// every launch that may mark frames gets its own pair, so launches of one plan on different streams do not share state
static int* next_redo_flags(const seld_plan* p) { return p->redo_pool + 2 * (__atomic_fetch_add(&p->redo_next, 1u, __ATOMIC_RELAXED) % kRedoPairs); }

static std::atomic<uint64_t> g_launches{0};
static thread_local int g_last_cuda = 0;

static int cuda_fail(cudaError_t e) {
    g_last_cuda = (int)e;
    return SELD_ECUDA;
}

extern "C" int seld_plan_create(seld_plan** out, int device, const float* window_host, const float* fb_host,
                                int n_fft, int hop, int n_mels, float amin, float eps) {
    if (!out || !window_host || !fb_host || hop <= 0 || n_mels <= 0) return SELD_EINVAL;
    if (n_fft != 1024) return SELD_EUNSUPPORTED;     // the warp-wide transform is 32 x 32
    const int F = n_fft / 2 + 1;

    // ---- band-sparse view of the mel bank: per band the [lo, lo+cnt) support and its weights
    std::vector<int> blo(n_mels), bcnt(n_mels), boff(n_mels);
    std::vector<float> wt;
    for (int m = 0; m < n_mels; ++m) {
        int lo = F, hi = -1;
        for (int k = 0; k < F; ++k)
            if (fb_host[(size_t)k * n_mels + m] != 0.0f) { if (k < lo) lo = k; hi = k; }
        blo[m] = hi < 0 ? 0 : lo;
        bcnt[m] = hi < 0 ? 0 : hi - lo + 1;
        boff[m] = (int)wt.size();
        for (int k = blo[m]; k < blo[m] + bcnt[m]; ++k) wt.push_back(fb_host[(size_t)k * n_mels + m]);
    }
    while (wt.size() % 4) wt.push_back(0.0f);
    if (wt.empty()) wt.assign(4, 0.0f);
    const int n_mels_pad = (n_mels + 3) & ~3;


    // ---- segment form for the iv2 kernel: bin k feeds only bands s_k-1 and s_k (true for every
    // triangular bank whose bands overlap their neighbours only); runs = maximal stretches of one
    // segment inside one lane's 16-bin chunk (bin 512 rides with lane 31).
    std::vector<float> wab(32 * 36, 0.0f);
    std::vector<uint32_t> runmask(32, 0);
    std::vector<int> g0(32, 0), gseg(n_mels + 2, 0);
    int fast_ok = (F == 513);
    // item form (see PlanDev): pieces of segments, one piece per lane and class
    std::vector<float> iw;
    std::vector<int> istart(4 * 32, 0);
    std::vector<uint32_t> islot((size_t)n_mels * 2, 0);
    int item_ok = 0, iP = 0, iK = 0, iL[4] = {0, 0, 0, 0}, ioff[4] = {0, 0, 0, 0};
    {
        std::vector<int> seg(F, 0);
        int sprev = 0;
        for (int k = 0; k < F && fast_ok; ++k) {
            int first = -1, last = -1, cnt = 0;
            for (int m = 0; m < n_mels; ++m)
                if (fb_host[(size_t)k * n_mels + m] != 0.0f) { if (first < 0) first = m; last = m; ++cnt; }
            int sk;
            if (cnt == 0) sk = sprev;
            else if (cnt == 1) {
                sk = (sprev == first || sprev == first + 1) ? sprev : first;
                // past the peak of a band with no upper neighbour here (the last band): its falling slope is the next segment
                if (sk == first && k > 0 && fb_host[(size_t)k * n_mels + first] < fb_host[(size_t)(k - 1) * n_mels + first]) sk = first + 1;
            }
            else if (cnt == 2 && last == first + 1) sk = last;
            else { fast_ok = 0; break; }
            if (sk < sprev) { fast_ok = 0; break; }
            seg[k] = sk; sprev = sk;
        }
        if (fast_ok) {
            int g = -1;
            std::vector<int> run_seg;
            for (int k = 0; k < F; ++k) {
                const int c = k < 512 ? k / 16 : 31, j = k - 16 * c;
                const bool start = (k == 0) || seg[k] != seg[k - 1] || (j == 0);
                if (start) { ++g; run_seg.push_back(seg[k]); if (j > 0) runmask[c] |= (1u << j); }
                if (j == 0) g0[c] = g;
                const int sk = seg[k];
                wab[c * 36 + 2 * j] = sk >= 1 ? fb_host[(size_t)k * n_mels + sk - 1] : 0.0f;
                wab[c * 36 + 2 * j + 1] = sk < n_mels ? fb_host[(size_t)k * n_mels + sk] : 0.0f;
            }
            const int nruns = g + 1;
            if (nruns > 127) fast_ok = 0;            // partial sums live in the first 254 words of a row; slot 127 stays zero
            for (int sgm = 0; sgm <= n_mels + 1; ++sgm) {
                int n = 0;
                for (int r = 0; r < nruns; ++r) if (run_seg[r] < sgm) ++n;
                gseg[sgm] = n;
            }
            for (int sgm = 0; sgm <= n_mels; ++sgm)
                if (gseg[sgm + 1] - gseg[sgm] > 4) fast_ok = 0;   // the combine step reads <= 4 runs per segment
        }
        if (fast_ok && n_mels <= 64) {
            // ---- item form.  Bins of segment s (those with a non-zero weight), cut into near-equal pieces of at most
            // `maxlen` bins; the 32 longest pieces form class 0, the next 32 class 1, ...; a class is as long as its longest
            // piece.  Choose (classes, maxlen) for the least work per frame: positions (each costs the per-bin arithmetic
            // and four loads) plus a per-class cost (its sums are stored and read back once).
            struct Piece { int lo, n, seg; };
            std::vector<int> slo(n_mels + 1, F), shi(n_mels + 1, -1);
            for (int k = 0; k < F; ++k) {
                const int sk = seg[k];
                const float wa = sk >= 1 ? fb_host[(size_t)k * n_mels + sk - 1] : 0.0f, wb = sk < n_mels ? fb_host[(size_t)k * n_mels + sk] : 0.0f;
                if (wa != 0.0f || wb != 0.0f) { if (k < slo[sk]) slo[sk] = k; shi[sk] = k; }
            }
            // Placement of one class: lane l reads bins start[l] + j, j < L, of the rows (natural bin order) with weight zero
            // outside its piece.  A warp-wide 64-bit load is two half-warp requests, each conflict-free when its 16 lanes start
            // at 16 different residues mod 16; a piece shorter than L may start up to L - n bins early, which is what makes the
            // residues assignable: maximum bipartite matching pieces <-> (half-warp, residue) = lane (Kuhn's augmenting paths).
            struct Place { int piece[32], start[32], wf; };
            auto place_class = [&](const std::vector<Piece>& cl, int L) {
                Place pl;
                for (int l = 0; l < 32; ++l) { pl.piece[l] = -1; pl.start[l] = l & 15; }   // idle lane: reads its own residue with zero weights
                std::vector<std::vector<int>> adj(cl.size());                              // adj[i]: starts piece i may take
                for (size_t i = 0; i < cl.size(); ++i)
                    for (int d = 0; d <= L - cl[i].n && d <= cl[i].lo; ++d)
                        if (cl[i].lo - d + L <= F) adj[i].push_back(cl[i].lo - d);         // a lane never reads past bin 512
                std::vector<int> start_of(cl.size(), -1);
                std::vector<char> seen;
                std::function<bool(int)> augment = [&](int i) -> bool {
                    for (int st : adj[i])
                        for (int h = 0; h < 2; ++h) {
                            const int l = 16 * h + (st & 15);
                            if (seen[l]) continue;
                            seen[l] = 1;
                            if (pl.piece[l] < 0 || augment(pl.piece[l])) { pl.piece[l] = i; pl.start[l] = st; start_of[i] = st; return true; }
                        }
                    return false;
                };
                std::vector<int> order(cl.size());
                for (size_t i = 0; i < cl.size(); ++i) order[i] = (int)i;
                std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return adj[x].size() < adj[y].size(); });
                std::vector<int> left;
                for (int i : order) { seen.assign(32, 0); if (adj[i].empty() || !augment(i)) left.push_back(i); }
                for (int i : left)                                          // no residue left for it: any free lane (those loads take an extra wavefront)
                    for (int l = 0; l < 32; ++l)
                        if (pl.piece[l] < 0) { pl.piece[l] = i; pl.start[l] = std::max(0, std::min(cl[i].lo, F - L)); break; }
                pl.wf = 0;
                for (int h = 0; h < 2; ++h) {
                    int cnt[16] = {}, mx = 0;
                    for (int l = 16 * h; l < 16 * h + 16; ++l) mx = std::max(mx, ++cnt[pl.start[l] & 15]);
                    pl.wf += mx;
                }
                return pl;
            };
            // A candidate plan = the pieces of every class and the class lengths (multiples of four: the kernel walks a class
            // four positions at a time).  Cost per position: 4 loads of `wf` wavefronts, the weights, the per-bin arithmetic;
            // per class: its sums are stored and read back once.
            struct Cand { std::vector<std::vector<Piece>> cls; std::vector<int> L; long cost = -1; };
            auto evaluate = [&](Cand& cd) {
                int P = 0; long cost = 28L * (long)cd.cls.size();
                for (size_t c = 0; c < cd.cls.size(); ++c) {
                    if (cd.cls[c].empty() || cd.cls[c].size() > 32) { cd.cost = -1; return; }
                    const Place pl = place_class(cd.cls[c], cd.L[c]);
                    cost += (long)cd.L[c] * (4 * pl.wf + 14);
                    P += cd.L[c];
                }
                cd.cost = (P >= 8 && P <= 24) ? cost : -1;
            };
            Cand best;
            const bool plan_debug = getenv("SELD_PLAN_DEBUG") && atoi(getenv("SELD_PLAN_DEBUG")) > 1;
            auto consider = [&](Cand& cd) {
                evaluate(cd);
                if (plan_debug && cd.cost >= 0) {
                    fprintf(stderr, "[seld plan] candidate (");
                    for (size_t c = 0; c < cd.L.size(); ++c) fprintf(stderr, "%s%d x%zu wf%d", c ? ", " : "", cd.L[c], cd.cls[c].size(), place_class(cd.cls[c], cd.L[c]).wf);
                    fprintf(stderr, ") cost %ld\n", cd.cost);
                }
                if (cd.cost >= 0 && (best.cost < 0 || cd.cost < best.cost)) best = cd;
            };
            // Near-equal cuts: every segment into ceil(n / maxlen) pieces, the 32 longest pieces form class 0, the next 32 class 1, ...
            // (Cutting to a tuple of class lengths instead -- (8, 4, 4, 4) = 20 positions fits the reference bank where this gives
            // (12, 8, 4) = 24 -- was measured 1.2 % SLOWER: a fourth class costs more in stored / re-read sums than four positions.)
            for (int K = 1; K <= 4; ++K)
                for (int maxlen = 1; maxlen <= 24; ++maxlen) {
                    std::vector<Piece> pc; bool ok = true;
                    for (int sgm = 0; sgm <= n_mels && ok; ++sgm) {
                        if (shi[sgm] < 0) continue;
                        const int n = shi[sgm] - slo[sgm] + 1, np = (n + maxlen - 1) / maxlen;
                        if (np > 4) { ok = false; break; }            // the combine step reads <= 4 pieces per segment
                        int at = slo[sgm];
                        for (int i = 0; i < np; ++i) { const int len = n / np + (i < n % np ? 1 : 0); pc.push_back({at, len, sgm}); at += len; }
                    }
                    if (!ok || (int)pc.size() > 32 * K || pc.empty()) continue;
                    std::stable_sort(pc.begin(), pc.end(), [](const Piece& x, const Piece& y) { return x.n > y.n; });
                    for (int slack = 0; slack <= 4; slack += 4) {
                        Cand cd;
                        for (int c = 0; c < K && (size_t)(32 * c) < pc.size(); ++c) {
                            cd.cls.emplace_back(pc.begin() + 32 * c, pc.begin() + std::min(pc.size(), (size_t)32 * (c + 1)));
                            cd.L.push_back((cd.cls.back()[0].n + slack + 3) & ~3);
                        }
                        consider(cd);
                    }
                }
            if (best.cost >= 0) {
                iK = (int)best.cls.size();
                for (int c = 0; c < 4; ++c) { iL[c] = c < iK ? best.L[c] : 0; ioff[c] = iP; iP += iL[c]; }
                {
                    item_ok = 1;
                    iw.assign((size_t)iP * 32 * 2, 0.0f);
                    std::vector<std::vector<int>> seg_slots(n_mels + 2);
                    for (int c = 0; c < iK; ++c) {
                        const std::vector<Piece>& cl = best.cls[c];
                        const Place pl = place_class(cl, iL[c]);
                        if (getenv("SELD_PLAN_DEBUG")) fprintf(stderr, "[seld plan] class %d: %d positions, %zu pieces, %d wavefronts per 64-bit load\n", c, iL[c], cl.size(), pl.wf);
                        for (int l = 0; l < 32; ++l) {
                            istart[c * 32 + l] = pl.start[l];
                            if (pl.piece[l] < 0) continue;
                            const Piece& pc = cl[pl.piece[l]];
                            seg_slots[pc.seg].push_back(32 * c + l);
                            for (int j = 0; j < pc.n; ++j) {
                                const int k = pc.lo + j, sk = seg[k], pp = ioff[c] + (k - pl.start[l]);
                                iw[((size_t)pp * 32 + l) * 2] = sk >= 1 ? fb_host[(size_t)k * n_mels + sk - 1] : 0.0f;
                                iw[((size_t)pp * 32 + l) * 2 + 1] = sk < n_mels ? fb_host[(size_t)k * n_mels + sk] : 0.0f;
                            }
                        }
                    }
                    for (int sgm = 0; sgm <= n_mels + 1; ++sgm) if (seg_slots[sgm].size() > 4) item_ok = 0;   // cannot happen: <= 4 pieces per segment
                    auto pack = [&](const std::vector<int>& v) {
                        uint32_t r = 0;
                        for (int i = 0; i < 4; ++i) r |= (uint32_t)(i < (int)v.size() ? v[i] : 128) << (8 * i);   // 128: the slot kept at zero
                        return r;
                    };
                    for (int m = 0; m < n_mels; ++m) { islot[2 * m] = pack(seg_slots[m]); islot[2 * m + 1] = pack(seg_slots[m + 1]); }
                }
            }
        }
    }
    if (iw.empty()) iw.assign(64, 0.0f);
    const int gseg_pad = (n_mels + 2 + 3) & ~3;

    // ---- twiddles W1024^(ka*j), window * 0.5
    std::vector<float> tw(2 * 1024), win(n_fft);
    for (int ka = 0; ka < 32; ++ka)
        for (int j = 0; j < 32; ++j) {
            const double ang = 2.0 * M_PI * (double)((ka * j) % 1024) / 1024.0;
            tw[2 * (ka * 32 + j)] = (float)cos(ang);
            tw[2 * (ka * 32 + j) + 1] = (float)(-sin(ang));
        }
    for (int i = 0; i < n_fft; ++i) win[i] = 0.5f * window_host[i];
    std::vector<float> tw4(16 * 32 * 4), win2(16 * 32 * 2);
    for (int q = 0; q < 16; ++q)
        for (int l = 0; l < 32; ++l) {
            const double a0 = 2.0 * M_PI * (double)((q * l) % 1024) / 1024.0;
            const double a1 = 2.0 * M_PI * (double)(((q + 16) * l) % 1024) / 1024.0;
            float* t = &tw4[(q * 32 + l) * 4];
            t[0] = (float)cos(a0); t[1] = (float)cos(a1); t[2] = (float)sin(a0); t[3] = (float)sin(a1);
            win2[(q * 32 + l) * 2] = win[32 * (2 * q) + l];
            win2[(q * 32 + l) * 2 + 1] = win[32 * (2 * q + 1) + l];
        }

    // ---- the bank as tensor-core B operand (iv5): 33 chunks of 16 bins, K-major unswizzled bf16 tiles holding the
    // bands each chunk touches, hi/lo halves interleaved along N (see MelTiles in seld_plan.h)
    std::vector<uint16_t> bimg;
    uint32_t mt_chunk[33] = {};
    int mt_ok = (F == 513 && n_mels == 64);
    if (mt_ok) {
        auto bf16_rn = [](float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16); };
        auto bf16_f = [](uint16_t h) { const uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; };
        for (int c = 0; c < 33 && mt_ok; ++c) {
            int lo = n_mels, hi = -1;
            for (int k = 16 * c; k < 16 * c + 16 && k < F; ++k)
                for (int m = 0; m < n_mels; ++m) {
                    const float w = fb_host[(size_t)k * n_mels + m];
                    if (!(fabsf(w) <= 3.0e38f)) mt_ok = 0;             // inf / nan weights: leave this form alone
                    if (w != 0.0f) { if (m < lo) lo = m; if (m > hi) hi = m; }
                }
            int col0 = 0, n = 8;
            if (hi >= 0) { col0 = lo & ~3; n = ((hi + 1 - col0) + 7) & ~7; if (col0 + n > n_mels) col0 = n_mels - n; }
            if (c == 0) { col0 = 0; n = n_mels; }                      // first chunk: full width, overwrites the accumulator
            const int N = 2 * n;
            const size_t off = bimg.size();                            // in uint16
            bimg.resize(off + (size_t)N * 16, 0);
            for (int kk = 0; kk < 16; ++kk) {
                const int k = 16 * c + kk;
                for (int j = 0; j < n; ++j) {
                    const float w = k < F ? fb_host[(size_t)k * n_mels + col0 + j] : 0.0f;
                    const uint16_t h = bf16_rn(w), l = bf16_rn(w - bf16_f(h));
                    for (int part = 0; part < 2; ++part) {
                        const int nn = 2 * j + part;
                        bimg[off + ((nn & 7) * 16 + (nn >> 3) * 256 + (kk & 7) * 2 + (kk >> 3) * 128) / 2] = part ? l : h;
                    }
                }
            }
            mt_chunk[c] = (uint32_t)(off * 2 / 16) | ((uint32_t)(2 * col0) << 16) | ((uint32_t)(N >> 3) << 24);
        }
        if (bimg.size() * 2 > 48 * 1024) mt_ok = 0;                    // dense banks: the tiles would not fit next to the rows
    }
    if (!mt_ok || bimg.empty()) bimg.assign(8, 0);

    seld_plan* p = new (std::nothrow) seld_plan();
    if (!p) return SELD_ENOMEM;
    memset(p, 0, sizeof(*p));
    p->device = device; p->n_fft = n_fft;

    int prev = 0;
    cudaError_t e = cudaGetDevice(&prev);
    if (e == cudaSuccess) e = cudaSetDevice(device);
    if (e != cudaSuccess) { delete p; return cuda_fail(e); }
    int smem_optin = 0;
    cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    p->smem_optin = (size_t)smem_optin;

    const size_t b_tw = tw.size() * 4, b_win = win.size() * 4, b_wt = wt.size() * 4, b_i = (size_t)n_mels_pad * 4;
    const size_t b_wab = wab.size() * 4, b_rm = 32 * 4, b_g0 = 32 * 4, b_gs = (size_t)gseg_pad * 4;
    const size_t b_tw4 = tw4.size() * 4, b_win2 = win2.size() * 4, b_bimg = bimg.size() * 2;
    const size_t b_iw = iw.size() * 4, b_idst = istart.size() * 4, b_islot = islot.size() * 4;
    const size_t total = b_tw + b_win + b_wt + 3 * b_i + b_wab + b_rm + b_g0 + b_gs + b_tw4 + b_win2 + b_bimg + 16 + b_iw + b_idst + b_islot + 32 + 16 + 2 * kRedoPairs * 4;
    e = cudaMalloc(&p->blob, total);
    if (e != cudaSuccess) { cudaSetDevice(prev); delete p; return cuda_fail(e); }
    std::vector<unsigned char> host(total, 0);
    size_t o = 0;
    memcpy(&host[o], tw.data(), b_tw); const size_t o_tw = o; o += b_tw;
    memcpy(&host[o], win.data(), b_win); const size_t o_win = o; o += b_win;
    memcpy(&host[o], wt.data(), b_wt); const size_t o_wt = o; o += b_wt;
    memcpy(&host[o], blo.data(), n_mels * 4); const size_t o_lo = o; o += b_i;
    memcpy(&host[o], bcnt.data(), n_mels * 4); const size_t o_cnt = o; o += b_i;
    memcpy(&host[o], boff.data(), n_mels * 4); const size_t o_off = o; o += b_i;
    memcpy(&host[o], wab.data(), b_wab); const size_t o_wab = o; o += b_wab;
    memcpy(&host[o], runmask.data(), b_rm); const size_t o_rm = o; o += b_rm;
    memcpy(&host[o], g0.data(), b_g0); const size_t o_g0 = o; o += b_g0;
    memcpy(&host[o], gseg.data(), gseg.size() * 4); const size_t o_gs = o; o += b_gs;
    memcpy(&host[o], tw4.data(), b_tw4); const size_t o_tw4 = o; o += b_tw4;
    memcpy(&host[o], win2.data(), b_win2); const size_t o_win2 = o; o += b_win2;
    o = (o + 15) & ~(size_t)15;                                        // the tile image is read as uint4
    memcpy(&host[o], bimg.data(), b_bimg); const size_t o_bimg = o; o += b_bimg;
    o = (o + 15) & ~(size_t)15;
    memcpy(&host[o], iw.data(), b_iw); const size_t o_iw = o; o += b_iw;
    memcpy(&host[o], istart.data(), b_idst); const size_t o_idst = o; o += b_idst;
    memcpy(&host[o], islot.data(), b_islot); const size_t o_islot = o; o += b_islot;
    o = (o + 15) & ~(size_t)15;
    const size_t o_redo = o; o += 2 * kRedoPairs * 4;                  // zeros
    e = cudaMemcpy(p->blob, host.data(), total, cudaMemcpyHostToDevice);
    cudaSetDevice(prev);
    if (e != cudaSuccess) { cudaFree(p->blob); delete p; return cuda_fail(e); }

    unsigned char* d = (unsigned char*)p->blob;
    p->dev.tw = (const float2*)(d + o_tw);
    p->dev.win = (const float*)(d + o_win);
    p->dev.tw4 = (const float4*)(d + o_tw4);
    p->dev.win2 = (const float2*)(d + o_win2);
    p->dev.wt = (const float*)(d + o_wt);
    p->dev.blo = (const int*)(d + o_lo);
    p->dev.bcnt = (const int*)(d + o_cnt);
    p->dev.boff = (const int*)(d + o_off);
    p->dev.wab = (const float*)(d + o_wab);
    p->dev.runmask = (const uint32_t*)(d + o_rm);
    p->dev.g0 = (const int*)(d + o_g0);
    p->dev.gseg = (const int*)(d + o_gs);
    p->mt.b_img = (const uint4*)(d + o_bimg);
    p->mt.b_bytes = (int)b_bimg; p->mt.ok = mt_ok;
    memcpy(p->mt.chunk, mt_chunk, sizeof(mt_chunk));
    p->dev.gseg_pad = gseg_pad;
    p->dev.fast_ok = fast_ok;
    p->redo_pool = (int*)(d + o_redo); p->redo_next = 0;
    p->dev.iw = (const float2*)(d + o_iw);
    p->dev.istart = (const int*)(d + o_idst);
    p->dev.islot = (const uint32_t*)(d + o_islot);
    p->dev.item_ok = item_ok; p->dev.iP = iP; p->dev.iK = iK;
    for (int c = 0; c < 4; ++c) { p->dev.iL[c] = iL[c]; p->dev.ioff[c] = ioff[c]; }
    p->dev.nnz_pad = (int)wt.size();
    p->dev.n_mels = n_mels; p->dev.n_mels_pad = n_mels_pad;
    p->dev.hop = hop; p->dev.amin = amin < 1.17549435e-38f ? 1.17549435e-38f : amin; p->dev.eps = eps;   // the kernels take log2 of max(v, amin) with a flush-to-zero MUFU: amin below FLT_MIN is raised to it

    // the tile (frames_per_tile-1)*hop + n_fft samples x 4 channels must fit next to the tables
    const int span = (((seld::foa_frames_per_tile() - 1) * hop + n_fft) + 3) & ~3;
    if (seld::foa_smem_bytes(p->dev, span) > p->smem_optin) { seld_plan_destroy(p); return SELD_EUNSUPPORTED; }
    p->use_iv2 = false; p->iv_kernel = 0;
    {
        // Kernel choice is fixed here, once per plan (nothing reads the environment on the per-call path).
        // SELD_IV_KERNEL is a developer switch for A/B runs: 1 = general kernel, 2 = fp32 mel walk (iv2, the default);
        // builds with -DSELD_EXPERIMENTS also know 5 = tensor-core mel projection (iv5: measured 0.478 ms against
        // 0.407 ms at cfg2) and 3 = two warps per frame (iv3).
        const char* force = getenv("SELD_IV_KERNEL");
        const int want = force ? atoi(force) : 2;
        p->iv2_ok = seld::foa_iv2_supported(p->dev, p->smem_optin);
        p->lm4_ok = want >= 2 && !getenv("SELD_NO_LM4") && seld::foa_iv2_supported(p->dev, p->smem_optin);
        if (false) {}
#ifdef SELD_EXPERIMENTS
        else if (want == 5 && seld::foa_iv5_supported(p->dev, p->mt, p->smem_optin)) p->iv_kernel = 5;
        else if (want == 3 && seld::foa_iv3_supported(p->dev, p->smem_optin)) p->iv_kernel = 3;
#endif
        else if (want >= 2 && seld::foa_iv2_supported(p->dev, p->smem_optin)) p->iv_kernel = 2;
        p->use_iv2 = p->iv_kernel != 0;
    }
    *out = p;
    return SELD_OK;
}

static void pipe_free_slots(HostPipe& hp) {
    for (int i = 0; i < HostPipe::kSlots; ++i) {
        if (hp.din[i]) cudaFree(hp.din[i]);
        if (hp.dout[i]) cudaFree(hp.dout[i]);
        hp.din[i] = hp.dout[i] = nullptr;
    }
    hp.cap_in = hp.cap_out = 0;
}

extern "C" void seld_plan_destroy(seld_plan* p) {
    if (!p) return;
    HostPipe& hp = p->pipe;
    if (hp.ready) {
        cudaStreamSynchronize(hp.s_in); cudaStreamSynchronize(hp.s_k); cudaStreamSynchronize(hp.s_out);
        pipe_free_slots(hp);
        cudaStreamDestroy(hp.s_in); cudaStreamDestroy(hp.s_k); cudaStreamDestroy(hp.s_out);
        cudaEventDestroy(hp.ev_start);
        for (int i = 0; i < HostPipe::kSlots; ++i) { cudaEventDestroy(hp.ev_in[i]); cudaEventDestroy(hp.ev_k[i]); cudaEventDestroy(hp.ev_out[i]); }
    }
    if (p->blob) cudaFree(p->blob);
    delete p;
}

extern "C" int64_t seld_num_frames(const seld_plan* p, int64_t L) {
    if (!p || L < 0) return SELD_EINVAL;
    return 1 + L / p->dev.hop;
}

static int run_foa(const seld_plan* p, bool iv, const void* x, int64_t B, int C, int64_t L,
                   int64_t stride_b, int64_t stride_c, float* out, void* stream, bool i16 = false) {
    if (!p || B < 0 || C < 1 || L < 1) return SELD_EINVAL;
    if (iv && C < 4) return SELD_EINVAL;               // intensityvector indexes channels 0..3
    if (L <= p->n_fft / 2) return SELD_ESHORT;         // reflect padding needs pad < L (torch.stft)
    if (B == 0) return SELD_OK;                        // empty batch: nothing to enqueue (pointers may be null)
    if (!x || !out) return SELD_EINVAL;
    const int64_t T = 1 + L / p->dev.hop;
    if (T > INT32_MAX) return SELD_EUNSUPPORTED;
    seld::FoaArgs a{};
    a.x = x; a.stride_b = stride_b; a.stride_c = stride_c; a.out = out; a.L = L;
    a.B = (int)B; a.C = C; a.Cout = C + (iv ? 3 : 0); a.T = (int)T; a.c_lo = 0;
    a.span = 0; a.vec_ok = 0; a.in_i16 = i16 ? 1 : 0; a.in_scale = i16 ? 1.0f / 32768.0f : 1.0f;
    if (i16 && !(iv && (p->iv_kernel == 2 || p->iv_kernel == 5) && C == 4)) return SELD_EUNSUPPORTED;   // PCM input: fused 4-channel path only
    cudaStream_t st = (cudaStream_t)stream;
    bool general_iv = iv;
    if (iv && p->use_iv2) {
        // channels 0-3: log-mel + IV by the packed dual-FFT kernel; any further channels below
        // iv5 stores 16-byte words: a misaligned output map goes to the fp32 mel walk instead
        int kern = p->iv_kernel;
        if (kern == 5 && ((uintptr_t)out & 15) != 0) {
            if (!p->iv2_ok) return SELD_EUNSUPPORTED;
            kern = 2;
        }
        int fpt = i16 ? 8 : seld::foa_iv2_frames_per_tile();          // PCM input always runs the 8-warp build
#ifdef SELD_EXPERIMENTS
        if (kern == 5) fpt = seld::foa_iv5_frames_per_tile();
        if (kern == 3) fpt = seld::foa_iv3_frames_per_tile();
#endif
        const int64_t tpc = (T + fpt - 1) / fpt;
        if (B * tpc > INT32_MAX) return SELD_EUNSUPPORTED;
        a.tiles_per_clip = (int)tpc; a.n_tiles = (int)(B * tpc);
        cudaError_t e;
        if (false) {}
#ifdef SELD_EXPERIMENTS
        else if (kern == 5) e = seld::foa_iv5_launch(a, p->dev, p->mt, p->sm_count, st);
        else if (kern == 3) e = seld::foa_iv3_launch(a, p->dev, p->sm_count, st);
#endif
        else { a.redo_flags = next_redo_flags(p); e = seld::foa_iv2_launch(a, p->dev, p->sm_count, st); }
        if (e != cudaSuccess) return cuda_fail(e);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (C == 4) return SELD_OK;
        a.c_lo = 4; general_iv = false;
    }
    if (!general_iv && p->lm4_ok && !i16) {
        // log-mel of channels [c_lo, C): the packed dual-FFT kernel in its log-mel-only mode
        const int64_t jobs = T * (int64_t)(C - a.c_lo);
        const int jpt = seld::foa_lm4_jobs_per_tile();
        const int64_t tpc = (jobs + jpt - 1) / jpt;
        if (B * tpc > INT32_MAX) return SELD_EUNSUPPORTED;
        a.tiles_per_clip = (int)tpc; a.n_tiles = (int)(B * tpc);
        a.redo_flags = next_redo_flags(p);
        cudaError_t e = seld::foa_lm4_launch(a, p->dev, p->sm_count, st);
        if (e != cudaSuccess) return cuda_fail(e);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return SELD_OK;
    }
    const int fpt = seld::foa_frames_per_tile();
    const int64_t tiles_per_clip = (T + fpt - 1) / fpt;
    if (B * tiles_per_clip > INT32_MAX) return SELD_EUNSUPPORTED;
    a.tiles_per_clip = (int)tiles_per_clip; a.n_tiles = (int)(B * tiles_per_clip);
    a.span = (((fpt - 1) * p->dev.hop + p->n_fft) + 3) & ~3;
    a.vec_ok = (((uintptr_t)x & 15) == 0) && (stride_b % 4 == 0) && (stride_c % 4 == 0) && (p->dev.hop % 4 == 0);
    cudaError_t e = seld::foa_launch(general_iv, a, p->dev, p->sm_count, st);
    if (e != cudaSuccess) return cuda_fail(e);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return SELD_OK;
}

extern "C" int seld_logmel_iv_f32(const seld_plan* p, const float* x, int64_t B, int C, int64_t L,
                                  int64_t stride_b, int64_t stride_c, float* out, void* stream) {
    return run_foa(p, true, x, B, C, L, stride_b, stride_c, out, stream);
}

extern "C" int seld_logmel_iv_i16(const seld_plan* p, const int16_t* x, int64_t B, int C, int64_t L,
                                  int64_t stride_b, int64_t stride_c, float* out, void* stream) {
    return run_foa(p, true, x, B, C, L, stride_b, stride_c, out, stream, true);
}

static int iv_host_pipeline(seld_plan* p, const void* x_host_v, bool i16, int64_t B, int C, int64_t L,
                            float* out_host, int chunk_clips, void* stream) {
    const unsigned char* x_host = (const unsigned char*)x_host_v;
    const size_t esz = i16 ? 2 : 4;
    if (!p || B < 0 || C < 4 || L < 1) return SELD_EINVAL;
    if (i16 && C != 4) return SELD_EUNSUPPORTED;
    if (L <= p->n_fft / 2) return SELD_ESHORT;
    if (B == 0) return SELD_OK;
    if (!x_host || !out_host) return SELD_EINVAL;
    const int64_t T = 1 + L / p->dev.hop;
    const int64_t in_clip = (int64_t)C * L, out_clip = (int64_t)(C + 3) * T * p->dev.n_mels;
    int64_t cc = chunk_clips > 0 ? chunk_clips : 4;         // measured on B200 + PCIe gen5: 2-4 clips per chunk is the sweet spot
    if (cc > B) cc = B;
    HostPipe& hp = p->pipe;
    int prev = 0;
    cudaError_t e = cudaGetDevice(&prev);
    if (e == cudaSuccess) e = cudaSetDevice(p->device);
    if (e != cudaSuccess) return cuda_fail(e);
#define SELD_TRY(call) do { e = (call); if (e != cudaSuccess) { cudaSetDevice(prev); return cuda_fail(e); } } while (0)
    if (!hp.ready) {
        SELD_TRY(cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking));
        SELD_TRY(cudaStreamCreateWithFlags(&hp.s_k, cudaStreamNonBlocking));
        SELD_TRY(cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking));
        SELD_TRY(cudaEventCreateWithFlags(&hp.ev_start, cudaEventDisableTiming));
        for (int i = 0; i < HostPipe::kSlots; ++i) {
            SELD_TRY(cudaEventCreateWithFlags(&hp.ev_in[i], cudaEventDisableTiming));
            SELD_TRY(cudaEventCreateWithFlags(&hp.ev_k[i], cudaEventDisableTiming));
            SELD_TRY(cudaEventCreateWithFlags(&hp.ev_out[i], cudaEventDisableTiming));
        }
        hp.ready = true;
    }
    const size_t need_in = (size_t)(cc * in_clip) * esz, need_out = (size_t)(cc * out_clip) * 4;
    if (need_in > hp.cap_in || need_out > hp.cap_out) {
        SELD_TRY(cudaStreamSynchronize(hp.s_in)); SELD_TRY(cudaStreamSynchronize(hp.s_k)); SELD_TRY(cudaStreamSynchronize(hp.s_out));
        pipe_free_slots(hp);
        for (int i = 0; i < HostPipe::kSlots; ++i) {
            SELD_TRY(cudaMalloc(&hp.din[i], need_in));
            SELD_TRY(cudaMalloc(&hp.dout[i], need_out));
        }
        hp.cap_in = need_in; hp.cap_out = need_out;
    }
    cudaStream_t user = (cudaStream_t)stream;
    SELD_TRY(cudaEventRecord(hp.ev_start, user));
    SELD_TRY(cudaStreamWaitEvent(hp.s_in, hp.ev_start, 0));
    SELD_TRY(cudaStreamWaitEvent(hp.s_k, hp.ev_start, 0));
    SELD_TRY(cudaStreamWaitEvent(hp.s_out, hp.ev_start, 0));
    int n = 0;
    for (int64_t b0 = 0; b0 < B; b0 += cc, ++n) {
        const int64_t nb = (B - b0 < cc) ? (B - b0) : cc;
        const int sl = n % HostPipe::kSlots;
        if (n >= HostPipe::kSlots) {                 // slot reuse: its previous kernel / copy-out must be done
            SELD_TRY(cudaStreamWaitEvent(hp.s_in, hp.ev_k[sl], 0));
            SELD_TRY(cudaStreamWaitEvent(hp.s_k, hp.ev_out[sl], 0));
        }
        SELD_TRY(cudaMemcpyAsync(hp.din[sl], x_host + (size_t)(b0 * in_clip) * esz, (size_t)(nb * in_clip) * esz, cudaMemcpyHostToDevice, hp.s_in));
        SELD_TRY(cudaEventRecord(hp.ev_in[sl], hp.s_in));
        SELD_TRY(cudaStreamWaitEvent(hp.s_k, hp.ev_in[sl], 0));
        const int rc = run_foa(p, true, hp.din[sl], nb, C, L, in_clip, L, hp.dout[sl], hp.s_k, i16);
        if (rc != SELD_OK) { cudaSetDevice(prev); return rc; }
        SELD_TRY(cudaEventRecord(hp.ev_k[sl], hp.s_k));
        SELD_TRY(cudaStreamWaitEvent(hp.s_out, hp.ev_k[sl], 0));
        SELD_TRY(cudaMemcpyAsync(out_host + b0 * out_clip, hp.dout[sl], (size_t)(nb * out_clip) * 4, cudaMemcpyDeviceToHost, hp.s_out));
        SELD_TRY(cudaEventRecord(hp.ev_out[sl], hp.s_out));
    }
    // later work on the caller's stream sees the finished batch
    for (int i = 0; i < HostPipe::kSlots && i < n; ++i) SELD_TRY(cudaStreamWaitEvent(user, hp.ev_out[i], 0));
#undef SELD_TRY
    cudaSetDevice(prev);
    return SELD_OK;
}

extern "C" int seld_logmel_iv_f32_host(seld_plan* p, const float* x_host, int64_t B, int C, int64_t L,
                                       float* out_host, int chunk_clips, void* stream) {
    return iv_host_pipeline(p, x_host, false, B, C, L, out_host, chunk_clips, stream);
}

extern "C" int seld_logmel_iv_i16_host(seld_plan* p, const int16_t* x_host, int64_t B, int C, int64_t L,
                                       float* out_host, int chunk_clips, void* stream) {
    return iv_host_pipeline(p, x_host, true, B, C, L, out_host, chunk_clips, stream);
}

extern "C" int seld_logmel_f32(const seld_plan* p, const float* x, int64_t B, int C, int64_t L,
                               int64_t stride_b, int64_t stride_c, float* out, void* stream) {
    return run_foa(p, false, x, B, C, L, stride_b, stride_c, out, stream);
}

extern "C" int64_t seld_num_frames_mic(const seld_plan* p, int64_t L) {
    if (!p || L < 0) return SELD_EINVAL;
    return L / p->dev.hop;
}

extern "C" size_t seld_workspace_bytes(const seld_plan* p, int64_t B, int C) {
    if (!p || B < 0 || C < 1) return 0;
    return 2 * (size_t)B * (size_t)C * sizeof(int);                     // per plane: running maximum and minimum
}

extern "C" int seld_logmel_gcc_f32(const seld_plan* p, const float* x, int64_t B, int C, int64_t L,
                                   int64_t stride_b, int64_t stride_c, float top_db, float* out,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    if (!p || B < 0 || C < 1 || L < 1) return SELD_EINVAL;
    if (C != 4) return SELD_EUNSUPPORTED;              // the CUDA path covers the 4-mic arrays of DCASE2021 / STARSS23
    if (!seld::mic_supported(p->dev, p->smem_optin)) return SELD_EUNSUPPORTED;
    const int64_t T = L / p->dev.hop;
    if (B == 0 || T == 0) return SELD_OK;
    if (!x || !out || !workspace) return SELD_EINVAL;
    if (workspace_bytes < seld_workspace_bytes(p, B, C)) return SELD_EINVAL;
    if (T > INT32_MAX) return SELD_EUNSUPPORTED;
    const int fpt = seld::mic_frames_per_tile();
    const int64_t tpc = (T + fpt - 1) / fpt;
    if (B * tpc > INT32_MAX) return SELD_EUNSUPPORTED;
    seld::FoaArgs a{};
    a.x = x; a.stride_b = stride_b; a.stride_c = stride_c; a.out = out; a.L = L;
    a.B = (int)B; a.C = C; a.Cout = C + C * (C - 1) / 2; a.T = (int)T; a.c_lo = 0;
    a.tiles_per_clip = (int)tpc; a.n_tiles = (int)(B * tpc); a.span = 0; a.vec_ok = 0; a.in_i16 = 0; a.in_scale = 1.0f;
    const bool use_top_db = top_db >= 0.0f;
    a.redo_flags = next_redo_flags(p);
    cudaError_t e = seld::mic_launch(a, p->dev, (int*)workspace, top_db, use_top_db, p->sm_count, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e);
    g_launches.fetch_add(use_top_db ? 2 : 1, std::memory_order_relaxed);
    return SELD_OK;
}

extern "C" int seld_mic_spectrogram_f32(const seld_plan* p, const float* x, int64_t B, int C, int64_t L,
                                        int64_t stride_b, int64_t stride_c, float* spec, void* stream) {
    if (!p || B < 0 || C < 1 || L < 1) return SELD_EINVAL;
    if (C != 4) return SELD_EUNSUPPORTED;
    if (!seld::mic_supported(p->dev, p->smem_optin)) return SELD_EUNSUPPORTED;
    const int64_t T = L / p->dev.hop;
    if (B == 0 || T == 0) return SELD_OK;
    if (!x || !spec) return SELD_EINVAL;
    if ((uintptr_t)spec & 15) return SELD_EUNSUPPORTED;
    if (T > INT32_MAX) return SELD_EUNSUPPORTED;
    const int fpt = seld::mic_frames_per_tile();
    const int64_t tpc = (T + fpt - 1) / fpt;
    if (B * tpc > INT32_MAX) return SELD_EUNSUPPORTED;
    seld::FoaArgs a{};
    a.x = x; a.stride_b = stride_b; a.stride_c = stride_c; a.out = nullptr; a.L = L; a.spec = spec;
    a.B = (int)B; a.C = C; a.Cout = C + C * (C - 1) / 2; a.T = (int)T; a.c_lo = 0;
    a.tiles_per_clip = (int)tpc; a.n_tiles = (int)(B * tpc); a.in_scale = 1.0f;
    cudaError_t e = seld::mic_spectrogram_launch(a, p->dev, p->sm_count, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return SELD_OK;
}

extern "C" int seld_logmel_gcc_from_spectra_f32(const seld_plan* p, const float* spec, int64_t B, int C, int64_t T,
                                                float top_db, float* out, void* workspace, size_t workspace_bytes,
                                                void* stream) {
    if (!p || B < 0 || C < 1 || T < 0) return SELD_EINVAL;
    if (C != 4) return SELD_EUNSUPPORTED;
    if (!seld::mic_supported(p->dev, p->smem_optin)) return SELD_EUNSUPPORTED;
    if (B == 0 || T == 0) return SELD_OK;
    if (!spec || !out || !workspace) return SELD_EINVAL;
    if ((uintptr_t)spec & 15) return SELD_EUNSUPPORTED;
    if (workspace_bytes < seld_workspace_bytes(p, B, C)) return SELD_EINVAL;
    if (T > INT32_MAX) return SELD_EUNSUPPORTED;
    const int fpt = seld::mic_frames_per_tile();
    const int64_t tpc = (T + fpt - 1) / fpt;
    if (B * tpc > INT32_MAX) return SELD_EUNSUPPORTED;
    seld::FoaArgs a{};
    a.x = nullptr; a.out = out; a.L = T * p->dev.hop; a.spec = const_cast<float*>(spec);
    a.B = (int)B; a.C = C; a.Cout = C + C * (C - 1) / 2; a.T = (int)T; a.c_lo = 0;
    a.tiles_per_clip = (int)tpc; a.n_tiles = (int)(B * tpc); a.in_scale = 1.0f;
    const bool use_top_db = top_db >= 0.0f;
    a.redo_flags = next_redo_flags(p);
    cudaError_t e = seld::mic_launch(a, p->dev, (int*)workspace, top_db, use_top_db, p->sm_count, (cudaStream_t)stream, true);
    if (e != cudaSuccess) return cuda_fail(e);
    g_launches.fetch_add(use_top_db ? 2 : 1, std::memory_order_relaxed);
    return SELD_OK;
}

// ---- backbone-input stage (stateless: no plan)
static int current_sm_count(int* sm) {
    static std::atomic<int> cache[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e);
    int v = dev < 64 ? cache[dev].load(std::memory_order_relaxed) : 0;
    if (v == 0) {
        e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return cuda_fail(e);
        if (dev < 64) cache[dev].store(v, std::memory_order_relaxed);
    }
    *sm = v;
    return SELD_OK;
}

static int check_scalar(const float* mean, const float* var, const float* weight, const float* bias, seld::ScalarArgs* s,
                        float eps) {
    if ((mean == nullptr) != (var == nullptr)) return SELD_EINVAL;      // statistics come as a pair
    if (!mean && (weight || bias)) return SELD_EINVAL;
    const uintptr_t al = (uintptr_t)mean | (uintptr_t)var | (uintptr_t)weight | (uintptr_t)bias;
    if (al & 15) return SELD_EUNSUPPORTED;
    s->mean = mean; s->var = var; s->weight = weight; s->bias = bias; s->eps = eps;
    return SELD_OK;
}

extern "C" int seld_scalar_f32(float* x, int64_t B, int C, int64_t T, int M, const float* mean, const float* var,
                               const float* weight, const float* bias, float eps, void* stream) {
    if (B < 0 || C < 1 || T < 1 || M < 1) return SELD_EINVAL;
    if (M % 4 != 0 || T > INT32_MAX) return SELD_EUNSUPPORTED;
    seld::ScalarArgs s;
    if (int rc = check_scalar(mean, var, weight, bias, &s, eps)) return rc;
    if (B == 0 || !mean) return SELD_OK;                                // nothing to do
    if (!x) return SELD_EINVAL;
    if ((uintptr_t)x & 15) return SELD_EUNSUPPORTED;
    int sm = 0;
    if (int rc = current_sm_count(&sm)) return rc;
    cudaError_t e = seld::scalar_launch(x, s, B, C, (int)T, M, sm, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return SELD_OK;
}

extern "C" int seld_scalar_wav2img_f32(const float* x, int64_t B, int C, int64_t T, int M, int spec_size,
                                       const float* mean, const float* var, const float* weight, const float* bias,
                                       float eps, float* img, void* stream) {
    if (B < 0 || C < 1 || T < 1 || M < 1 || spec_size < 1) return SELD_EINVAL;
    if (spec_size % M != 0) return SELD_EINVAL;                         // htsat.py:442,506: the fold needs spec_size = r * M
    if (M % 4 != 0 || spec_size % 4 != 0 || T > INT32_MAX) return SELD_EUNSUPPORTED;
    seld::ScalarArgs s;
    if (int rc = check_scalar(mean, var, weight, bias, &s, eps)) return rc;
    if (B == 0) return SELD_OK;
    if (!x || !img) return SELD_EINVAL;
    if (((uintptr_t)x | (uintptr_t)img) & 15) return SELD_EUNSUPPORTED;
    int sm = 0;
    if (int rc = current_sm_count(&sm)) return rc;
    cudaError_t e = seld::scalar_wav2img_launch(x, img, s, B, C, (int)T, M, spec_size, sm, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return SELD_OK;
}

// ---- waveform-domain augmentation
extern "C" int seld_foa_rotate_f32(float* x, int64_t B, int C, int64_t L, int64_t stride_b, int64_t stride_c,
                                   const int32_t* codes, void* stream) {
    if (B < 0 || C < 4 || L < 1) return SELD_EINVAL;                    // rotate.py indexes channels 0..3
    if (B == 0) return SELD_OK;
    if (!x || !codes) return SELD_EINVAL;
    if ((uintptr_t)x & 3) return SELD_EINVAL;
    cudaError_t e = seld::foa_rotate_launch(x, B, L, stride_b, stride_c, codes, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return SELD_OK;
}

extern "C" int seld_wavmix_order(const int64_t* dst, const int64_t* src, const float* lam, int n, int64_t B,
                                 seld_mix_op* out) {
    if (n < 0 || B < 0) return SELD_EINVAL;
    if (n == 0) return SELD_OK;
    if (!dst || !src || !lam || !out) return SELD_EINVAL;
    std::vector<int> writer(B, -1), reader(B, -1);                      // op that writes / reads clip b
    for (int k = 0; k < n; ++k) {
        if (dst[k] < 0 || dst[k] >= B || src[k] < 0 || src[k] >= B) return SELD_EINVAL;
        if (writer[dst[k]] >= 0 || reader[src[k]] >= 0) return SELD_EINVAL;   // repeated index
        writer[dst[k]] = k; reader[src[k]] = k;
    }
    std::vector<char> done(n, 0);
    int m = 0;
    auto emit = [&](int k, int flags) {
        out[m].dst = (int32_t)dst[k]; out[m].src = (int32_t)src[k]; out[m].lam = lam[k]; out[m].flags = flags;
        done[k] = 1; ++m;
    };
    // open chains: start at a destination nobody reads, follow dst -> src while the source is itself mixed
    for (int k0 = 0; k0 < n; ++k0) {
        if (done[k0] || reader[dst[k0]] >= 0) continue;
        int k = k0, flags = SELD_MIX_BEGIN;
        while (k >= 0 && !done[k]) { emit(k, flags); flags = 0; k = writer[src[k]]; }
    }
    // what is left are closed cycles: the last op's source is the (overwritten) first destination
    for (int k0 = 0; k0 < n; ++k0) {
        if (done[k0]) continue;
        int k = k0, flags = SELD_MIX_BEGIN;
        while (!done[k]) {
            const bool closes = src[k] == dst[k0];
            emit(k, flags | (closes ? SELD_MIX_USE_HEAD : 0));
            flags = 0;
            if (closes) break;
            k = writer[src[k]];
        }
    }
    return m == n ? SELD_OK : SELD_EINVAL;
}

extern "C" int seld_wavmix_f32(float* x, int64_t B, int C, int64_t L, int64_t stride_b, int64_t stride_c,
                               const seld_mix_op* ops, int n_ops, void* stream) {
    if (B < 0 || C < 1 || L < 1 || n_ops < 0) return SELD_EINVAL;
    if (B == 0 || n_ops == 0) return SELD_OK;
    if (!x || !ops) return SELD_EINVAL;
    if (((uintptr_t)x & 3) || ((uintptr_t)ops & 15)) return SELD_EINVAL;
    if (C > 65535) return SELD_EUNSUPPORTED;
    cudaError_t e = seld::wavmix_launch(x, C, L, stride_b, stride_c, ops, n_ops, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return SELD_OK;
}

extern "C" uint64_t seld_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" int seld_last_cuda_error(void) { return g_last_cuda; }

extern "C" const char* seld_strerror(int code) {
    switch (code) {
        case SELD_OK: return "ok";
        case SELD_EINVAL: return "invalid argument";
        case SELD_EUNSUPPORTED: return "unsupported configuration (n_fft must be 1024; tile must fit shared memory)";
        case SELD_ESHORT: return "clip too short: reflect padding needs n_fft/2 < L";
        case SELD_ECUDA: return "CUDA runtime error";
        case SELD_ENOMEM: return "out of memory";
        default: return "unknown error";
    }
}

extern "C" const char* seld_version(void) { return "seldfeat 0.1 sm_100a"; }
