// expand_fused.cu -- cnt2event with linear timestamps in three launches, no host round trip and no per-event intermediate.
//
// Replaces, for the case the pipeline runs every step (dataloader/cython_cnt2event/cnt2event.pyx:18-116 with mode 'linear',
// per-pixel counts <= 64), the count -> scan -> emit items -> histogram -> scan -> scatter chain of events.cu.  Observations:
//   * the timestamp of event j of a pixel holding n events is float32(np.linspace(0, 1, n)[j]): one of K(m) distinct values when
//     all counts are <= m (K(8) = 19, K(64) = 1229), so the per-sample stable sort by time is ONE counting sort over K keys;
//   * how many events of key k precede a tile follows from how many pixels of each count n precede it (a pixel of count n emits
//     key k at most once), so the global structure is a per-tile histogram over n (66 bins), not over keys or events;
//   * the padded [B, maxlen, 4] layout needs no cross-sample scan: row = b * maxlen + (events of the sample with a smaller key)
//     + (events of this key in earlier pixels).  maxlen is computed on the device by every CTA from the per-sample statistics;
//     the caller supplies a row capacity and reads the statistics afterwards (one synchronisation, at the end).
// Passes: k_xf_count (round, per-tile count histogram, statistics), k_xf_scan (exclusive scan of each (sample, n) row over
// tiles; the last block of a sample turns the row totals into the per-key row offsets of the sample and publishes the sizing
// decision), k_xf_emit (re-reads the values, ranks its events inside the tile in emission order with match.any, writes final rows
// and the zero padding).  A tile is the 256 consecutive pixel slots of ONE warp: the warps of a CTA never synchronise with each
// other.  HBM traffic: 2 x 4 B per pixel slot + 16 B per output row (+ 264 B of histogram per tile).
//
// Anything this path cannot do (negative counts in an active sample, a count > 64, capacity too small) is detected from the same
// statistics by the host, which then runs the general chain of events.cu -- on the GPU as well; there is no CPU path.
#include "common.cuh"

namespace esr {

constexpr int XF_WARPS = 8;
constexpr int XF_WSLOTS = 256;                 // pixel slots per warp = one tile
constexpr int XF_TILE = XF_WARPS * XF_WSLOTS;  // 2048 slots per CTA
constexpr int XF_EVMAP = 1024;                 // tiles with at most this many events resolve event -> slot through a table
constexpr int XF_MAXN = 64;                    // largest per-slot count handled here
constexpr int XF_NBINS = XF_MAXN + 2;          // counts 0..64, bin 65 = anything larger
constexpr int XF_NTAB = 7;                     // key tables for m = 1, 2, 4, ..., 64
constexpr int XF_MAXK = 1232;                  // K(64) = 1229 distinct timestamps, padded

struct XfTables {
    const uint16_t *rank[XF_NTAB];             // rank[i][n * m + j], m = 1 << i
    const float *uniq[XF_NTAB];                // the K[i] distinct timestamps, ascending
    int K[XF_NTAB];
};

__device__ __forceinline__ unsigned int lanemask_lt()
{
    unsigned int m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// rounded value -> per-slot event count, as k_expand_count (events.cu) for kind 0.  cvt.rni.s32.f32 rounds half to even like
// numpy round() (cnt2event.pyx:31) and saturates; |values| >= 2^31 are far outside the fused path (max count 64) either way.
struct XfAcc { long long sum = 0; unsigned long long nev = 0; int neg = 0; unsigned int mx = 0; };
__device__ __forceinline__ unsigned int xf_count(float x, XfAcc &a)
{
    const int ri = __float2int_rn(x);
    a.sum += ri;
    a.neg |= ri < 0;
    const unsigned int n = (unsigned int)max(ri, 0);
    a.nev += n;
    a.mx = max(a.mx, n);
    return n;
}
__device__ __forceinline__ unsigned int xf_n(float x) { return (unsigned int)min(max(__float2int_rn(x), 0), XF_MAXN); }

// 4 consecutive slots of lane `lane` in half-chunk g of warp w (slot order inside a warp: g, lane, element)
template <bool VEC>
__device__ __forceinline__ void xf_load4(const float *__restrict__ v, int64_t rem, int idx, float x[4])
{
    if (VEC && idx + 3 < rem) {
        const float4 q = __ldg(reinterpret_cast<const float4 *>(v + idx));
        x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) x[e] = (idx + e < rem) ? __ldg(v + idx + e) : 0.0f;
    }
}

// ---------------------------------------------------------------------------------------------
// pass 1: per-tile (= per-warp) histogram of the per-slot counts + per-sample statistics.  grid = (ceil(Tw / 8), B)
// ---------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(XF_WARPS * 32)
k_xf_count(const float *__restrict__ vals, int64_t S, int Tw, uint16_t *__restrict__ hn16 /*[B][Tw][XF_NBINS]: slots per count, per tile*/,
           unsigned long long *__restrict__ stats /*[B][4]: sum, events, any negative, max count*/)
{
    PDL_LAUNCH_DEPENDENTS();
    __shared__ uint32_t hw[XF_WARPS][XF_NBINS + 30];
    __shared__ long long s_sum[XF_WARPS];
    __shared__ unsigned long long s_nev[XF_WARPS];
    __shared__ int s_neg[XF_WARPS];
    __shared__ unsigned int s_mx[XF_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k = lane; k < XF_NBINS + 30; k += 32) hw[w][k] = 0;
    __syncwarp();
    PDL_WAIT();
    const int b = blockIdx.y, wt = blockIdx.x * XF_WARPS + w;
    const float *v = vals + (size_t)b * S + (size_t)wt * XF_WSLOTS;
    const int64_t rem = S - (int64_t)wt * XF_WSLOTS;
    XfAcc a;
    float x[2][4];
#pragma unroll
    for (int g = 0; g < 2; ++g) xf_load4<VEC>(v, rem, g * 128 + lane * 4, x[g]);
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const unsigned int n = xf_count(x[g][e], a);
            const unsigned int bin = min(n, (unsigned int)(XF_MAXN + 1));
            const unsigned int peers = __match_any_sync(0xffffffffu, bin);
            if (lane == __ffs(peers) - 1) hw[w][bin] += __popc(peers);
            __syncwarp();
        }
    if (wt < Tw) {                                          // 132 contiguous bytes per tile (a tile has 256 slots: counts fit 16 bits)
        uint32_t *dst = reinterpret_cast<uint32_t *>(hn16 + ((size_t)b * Tw + wt) * XF_NBINS);
        dst[lane] = hw[w][2 * lane] | (hw[w][2 * lane + 1] << 16);
        if (lane == 0) dst[32] = hw[w][64] | (hw[w][65] << 16);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        a.sum += __shfl_xor_sync(0xffffffffu, a.sum, d);
        a.nev += __shfl_xor_sync(0xffffffffu, a.nev, d);
        a.neg |= __shfl_xor_sync(0xffffffffu, a.neg, d);
        a.mx = max(a.mx, __shfl_xor_sync(0xffffffffu, a.mx, d));
    }
    if (lane == 0) { s_sum[w] = a.sum; s_nev[w] = a.nev; s_neg[w] = a.neg; s_mx[w] = a.mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int ww = 1; ww < XF_WARPS; ++ww) { a.sum += s_sum[ww]; a.nev += s_nev[ww]; a.neg |= s_neg[ww]; a.mx = max(a.mx, s_mx[ww]); }
        unsigned long long *st = stats + (size_t)b * 4;
        if (a.sum != 0) atomicAdd(st + 0, (unsigned long long)a.sum);      // two's complement add == signed add
        if (a.nev != 0) atomicAdd(st + 1, a.nev);
        if (a.neg) atomicOr(st + 2, 1ull);
        if (a.mx) atomicMax(st + 3, (unsigned long long)a.mx);
    }
}

// ---------------------------------------------------------------------------------------------
// pass 2: sizing decision; exclusive scan of every (sample, count) row over the tiles, in place; per-key row offsets of each
// sample (by the last block of the sample to finish).  grid = (largest count the caller allows + 1, B)
// ---------------------------------------------------------------------------------------------
struct XfMeta { unsigned long long maxlen; unsigned int mx, status, ti, pad; };   // status 0: the rows are being written

__global__ void __launch_bounds__(256)
k_xf_scan(const uint16_t *__restrict__ hn16, uint32_t *__restrict__ pre /*[B][XF_NBINS][Tw]: slots of count n in earlier tiles*/, int Tw, int B,
          const unsigned long long *__restrict__ stats, const XfTables tab, int ti_cap,
          unsigned long long cap_rows, uint32_t *__restrict__ tot /*[B][XF_NBINS]*/, uint32_t *__restrict__ keybase /*[B][XF_MAXK]*/,
          unsigned int *__restrict__ done /*[B], zero*/, XfMeta *__restrict__ meta)
{
    PDL_LAUNCH_DEPENDENTS();
    __shared__ uint32_t ws[8];
    __shared__ uint32_t s_carry;
    __shared__ unsigned long long s_maxlen;
    __shared__ unsigned int s_mx, s_last;
    __shared__ int s_err;
    __shared__ uint32_t s_key[XF_MAXK];
    if (threadIdx.x == 0) { s_carry = 0; s_maxlen = 1; s_mx = 0; s_err = 0; }
    __syncthreads();
    PDL_WAIT();
    // the reference's sizing rules, from the statistics (cnt2event.pyx:33-60)
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const unsigned long long *st = stats + (size_t)i * 4;
        if (st[0] != 0) {                                   // rounded values sum to non-zero: the sample emits its events
            atomicMax(&s_maxlen, st[1]);
            atomicMax(&s_mx, (unsigned int)min(st[3], 0xffffffffull));
            if (st[2]) s_err = 1;                           // np.zeros([negative, 4]) raises in the reference
        }
    }
    __syncthreads();
    const unsigned long long maxlen = s_maxlen;
    const unsigned int mx = s_mx;
    const unsigned int status = s_err ? 1u : mx == 0 ? 4u : mx > (1u << ti_cap) ? 2u : maxlen * (unsigned long long)B > cap_rows ? 3u : 0u;
    const int n = blockIdx.x, b = blockIdx.y;
    const bool active = stats[(size_t)b * 4] != 0;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (status == 0 && active && n >= 1 && n <= (int)mx) {
        uint32_t *row = pre + ((size_t)b * XF_NBINS + n) * Tw;
        const uint16_t *src = hn16 + (size_t)b * Tw * XF_NBINS + n;
        for (int base = 0; base < Tw; base += 256) {
            const int i = base + threadIdx.x;
            const uint32_t v = i < Tw ? (uint32_t)src[(size_t)i * XF_NBINS] : 0u;
            uint32_t incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += u;
            }
            if (lane == 31) ws[w] = incl;
            __syncthreads();
            uint32_t wpre = 0, all = 0;
#pragma unroll
            for (int ww = 0; ww < 8; ++ww) { if (ww < w) wpre += ws[ww]; all += ws[ww]; }
            const uint32_t carry = s_carry;
            if (i < Tw) row[i] = carry + wpre + incl - v;
            __syncthreads();
            if (threadIdx.x == 0) s_carry = carry + all;
            __syncthreads();
        }
        if (threadIdx.x == 0) tot[(size_t)b * XF_NBINS + n] = s_carry;
    }
    // last block of the sample: per-key row offsets
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(done + b, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    int ti = 0;
    while ((1u << ti) < mx) ++ti;
    if (b == 0 && threadIdx.x == 0) { meta->maxlen = maxlen; meta->mx = mx; meta->status = status; meta->ti = (unsigned int)ti; }
    if (status != 0 || !active) return;
    const int m = 1 << ti, K = tab.K[ti];
    const uint16_t *__restrict__ rank = tab.rank[ti];
    for (int k = threadIdx.x; k < K; k += blockDim.x) s_key[k] = 0;
    __syncthreads();
    for (int idx = threadIdx.x; idx < (int)(mx + 1) * m; idx += blockDim.x) {
        const int nn = idx >> ti, j = idx & (m - 1);
        if (nn == 0 || j >= nn) continue;
        const uint32_t c = __ldcg(tot + (size_t)b * XF_NBINS + nn);      // written by the other blocks of this sample
        if (c) atomicAdd(&s_key[rank[idx]], c);
    }
    __syncthreads();
    {   // exclusive scan over the K keys: thread i owns keys [i*per, i*per+per)
        const int per = (K + (int)blockDim.x - 1) / (int)blockDim.x;
        const int k0 = threadIdx.x * per;
        uint32_t mine = 0;
        for (int q = 0; q < per; ++q) if (k0 + q < K) mine += s_key[k0 + q];
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += u;
        }
        if (lane == 31) ws[w] = incl;
        __syncthreads();
        uint32_t run = incl - mine;
        for (int ww = 0; ww < w; ++ww) run += ws[ww];
        for (int q = 0; q < per; ++q)
            if (k0 + q < K) { keybase[(size_t)b * XF_MAXK + k0 + q] = run; run += s_key[k0 + q]; }
    }
}

// ---------------------------------------------------------------------------------------------
// pass 3: final rows.  grid = (ceil(Tw / 8), B); dynamic shared memory: XF_WARPS x (XfWarpSmem + kcap counters)
// ---------------------------------------------------------------------------------------------
struct __align__(16) XfWarpSmem {
    uint32_t pre_n[XF_NBINS + 2];        // slots of count n in earlier tiles of the sample
    uint16_t P[XF_WSLOTS + 2];           // exclusive prefix of the counts inside the tile, [256] = total
    uint16_t evmap[XF_EVMAP];            // event -> slot << 6 | j   (tiles with <= XF_EVMAP events)
    uint8_t nS[XF_WSLOTS];
};
static size_t xf_emit_smem(int kcap) { return (size_t)XF_WARPS * (sizeof(XfWarpSmem) + sizeof(uint32_t) * (size_t)kcap); }

template <bool VEC>
__global__ void __launch_bounds__(XF_WARPS * 32)
k_xf_emit(const float *__restrict__ vals, int64_t S, int Tw, int H, int W, const uint32_t *__restrict__ pre,
          const uint32_t *__restrict__ keybase, const unsigned long long *__restrict__ stats, const XfMeta *__restrict__ meta,
          const XfTables tab, int kcap, float *__restrict__ out)
{
    PDL_LAUNCH_DEPENDENTS();
    extern __shared__ __align__(16) unsigned char xf_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    XfWarpSmem &sm = reinterpret_cast<XfWarpSmem *>(xf_raw)[w];
    uint32_t *const bw = reinterpret_cast<uint32_t *>(xf_raw + XF_WARPS * sizeof(XfWarpSmem)) + (size_t)w * kcap;   // next row per key
    const int b = blockIdx.y, wt = blockIdx.x * XF_WARPS + w;
    PDL_WAIT();
    // this tile's values are on their way while the sizing decision is fetched
    const float *v = vals + (size_t)b * S + (size_t)wt * XF_WSLOTS;
    const int64_t rem = S - (int64_t)wt * XF_WSLOTS;
    float x[2][4];
#pragma unroll
    for (int g = 0; g < 2; ++g) xf_load4<VEC>(v, rem, g * 128 + lane * 4, x[g]);
    if (meta->status != 0) return;                          // the host sees the same statistics and takes the general chain
    const unsigned long long maxlen = meta->maxlen;
    const unsigned int mx = meta->mx;
    const int ti = (int)meta->ti;
    const unsigned long long *myst = stats + (size_t)b * 4;
    const bool active = myst[0] != 0;
    const unsigned long long ev = active ? myst[1] : 0ull;
    const size_t row0 = (size_t)b * maxlen;
    float4 *o = reinterpret_cast<float4 *>(out) + row0;
    {   // zero padding behind the sample's events (the whole sample when it is inactive): this CTA's share
        const uint32_t nrows = (uint32_t)(maxlen - ev), chunk = nrows / gridDim.x + 1u;       // B * maxlen < 2^32 (cap_rows)
        const uint32_t lo = min(nrows, chunk * blockIdx.x), hi = min(nrows, lo + chunk);
        float4 *pad = o + ev;
        for (uint32_t r = lo + threadIdx.x; r < hi; r += blockDim.x) pad[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (!active || wt >= Tw) return;

    const int m = 1 << ti, K = tab.K[ti];
    const uint16_t *__restrict__ rank = tab.rank[ti];
    const float *__restrict__ uniq = tab.uniq[ti];
    // rows of this tile's keys: per-sample key offset + events of the key in earlier tiles
    for (int k = lane; k < K; k += 32) bw[k] = __ldg(keybase + (size_t)b * XF_MAXK + k);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const int n = q * 32 + lane;
        if (n >= 1 && n <= (int)mx) sm.pre_n[n] = __ldg(pre + ((size_t)b * XF_NBINS + n) * Tw + wt);
    }
    // counts, exclusive prefix inside the tile, event -> (slot, j) table
    uint32_t E;
    {
        uint32_t carry = 0;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const uint32_t n0 = xf_n(x[g][0]), n1 = xf_n(x[g][1]), n2 = xf_n(x[g][2]), n3 = xf_n(x[g][3]);
            const uint32_t c1 = n0, c2 = c1 + n1, c3 = c2 + n2, mine = c3 + n3;
            uint32_t incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += u;
            }
            const uint32_t run = carry + incl - mine;
            const int s0 = g * 128 + lane * 4;
            const uint32_t tot_g = __shfl_sync(0xffffffffu, incl, 31);
            *reinterpret_cast<uint2 *>(&sm.P[s0]) = make_uint2(run | ((run + c1) << 16), (run + c2) | ((run + c3) << 16));
            *reinterpret_cast<uint32_t *>(&sm.nS[s0]) = n0 | (n1 << 8) | (n2 << 16) | (n3 << 24);
            if (carry + tot_g <= (uint32_t)XF_EVMAP)
                for (uint32_t q = 0; q < mine; ++q) {
                    const uint32_t e = (q >= c1) + (q >= c2) + (q >= c3);
                    const uint32_t first = e == 0 ? 0u : e == 1 ? c1 : e == 2 ? c2 : c3;
                    sm.evmap[run + q] = (uint16_t)(((s0 + e) << 6) | (q - first));
                }
            carry += tot_g;
        }
        E = carry;
        if (lane == 0) sm.P[XF_WSLOTS] = (uint16_t)E;
    }
    __syncwarp();
    // a slot of count n emits key rank[n][j] once for every j < n: earlier tiles' count histogram -> key offsets
    for (int idx = lane; idx < (int)(mx + 1) * m; idx += 32) {
        const int nn = idx >> ti, j = idx & (m - 1);
        if (nn == 0 || j >= nn) continue;
        const uint32_t c = sm.pre_n[nn];
        if (c) atomicAdd(&bw[rank[idx]], c);
    }
    __syncwarp();
    // ---- the warp walks its events in emission order (slot-major, then j), 32 per round; equal keys keep that order
    const uint32_t HW = (uint32_t)H * (uint32_t)W;
    const uint32_t in0 = (uint32_t)wt * XF_WSLOTS;
    const uint32_t pol0 = in0 >= HW ? 1u : 0u, in_r = in0 - pol0 * HW;                        // first slot of the tile
    const uint32_t y0 = in_r / (uint32_t)W, x0 = in_r - y0 * (uint32_t)W;
    const bool wide = W >= XF_WSLOTS;                                                          // at most one row wrap inside a tile
    const unsigned int lt = lanemask_lt();
    const bool use_map = E <= (uint32_t)XF_EVMAP;
    for (uint32_t e0 = 0; e0 < E; e0 += 32) {
        const uint32_t e = e0 + lane;
        const bool valid = e < E;
        uint32_t s = 0, j = 0;
        if (use_map) {
            const uint32_t q = valid ? sm.evmap[e] : 0u;
            s = q >> 6; j = q & 63u;
        } else {
#pragma unroll
            for (int step = 128; step > 0; step >>= 1)
                if (sm.P[s + step] <= e) s += step;        // last slot whose prefix is <= e (empty slots share their successor's prefix)
            j = e - sm.P[s];
        }
        uint32_t k = 0xffffffffu;
        if (valid) k = rank[((uint32_t)sm.nS[s] << ti) + j];
        const unsigned int peers = __match_any_sync(0xffffffffu, k);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && lane == leader) { old = bw[k]; bw[k] = old + __popc(peers); }
        old = __shfl_sync(0xffffffffu, old, leader);
        __syncwarp();
        if (valid) {
            uint32_t xx, yy, pol;
            if (wide) {
                xx = x0 + s; yy = y0; pol = pol0;
                if (xx >= (uint32_t)W) { xx -= (uint32_t)W; ++yy; }
                if (yy >= (uint32_t)H) { yy -= (uint32_t)H; ++pol; }
            } else {
                const uint32_t in_s = in0 + s;
                xx = in_s % (uint32_t)W; yy = (in_s / (uint32_t)W) % (uint32_t)H; pol = in_s / HW;
            }
            float4 row;
            row.x = (float)xx; row.y = (float)yy;
            row.z = __ldg(uniq + k);
            row.w = (pol & 1u) ? -1.0f : 1.0f;                       // channel 1 = negative polarity (cnt2event.pyx:80-90)
            o[old + __popc(peers & lt)] = row;
        }
    }
}

} // namespace esr

using namespace esr;

static size_t xf_ws_layout(int B, int Tw, size_t *o_hn, size_t *o_pre, size_t *o_tot, size_t *o_key, size_t *o_done, size_t *o_meta)
{
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t r = off; off = (off + bytes + 255) & ~(size_t)255; return r; };
    *o_done = take(sizeof(unsigned int) * (size_t)B);
    *o_meta = take(sizeof(XfMeta));
    *o_hn = take(sizeof(uint16_t) * (size_t)B * XF_NBINS * (size_t)Tw);
    *o_pre = take(sizeof(uint32_t) * (size_t)B * XF_NBINS * (size_t)Tw);
    *o_tot = take(sizeof(uint32_t) * (size_t)B * XF_NBINS);
    *o_key = take(sizeof(uint32_t) * (size_t)B * XF_MAXK);
    return off;
}

extern "C" size_t esr_cnt2event_fused_workspace_bytes(int B, int H, int W)
{
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    const int64_t S = 2ll * H * W;
    size_t a, b, c, d, e, f;
    return xf_ws_layout(B, (int)((S + XF_WSLOTS - 1) / XF_WSLOTS), &a, &b, &c, &d, &e, &f);
}

extern "C" int esr_cnt2event_fused(const float *vals, int B, int H, int W, const void *tables, const int32_t *tables_desc_host,
                                   int max_count, int64_t *stats, float *out, int64_t cap_rows, void *workspace,
                                   size_t workspace_bytes, esr_stream_t stream)
{
    ESR_REQUIRE(vals && tables && tables_desc_host && stats && out && workspace, "esr_cnt2event_fused: null pointer");
    ESR_REQUIRE(B > 0 && B <= 256 && H > 0 && W > 0 && cap_rows > 0, "esr_cnt2event_fused: bad dims B=%d H=%d W=%d cap=%lld", B, H, W,
                (long long)cap_rows);
    const int64_t S = 2ll * H * W;
    ESR_REQUIRE((int64_t)B * S < (1ll << 32) && S < (1ll << 31), "esr_cnt2event_fused: more than 2^32 slots");
    ESR_REQUIRE(cap_rows < (1ll << 32), "esr_cnt2event_fused: row capacity must stay below 2^32");
    ESR_REQUIRE(max_count >= 1 && max_count <= XF_MAXN, "esr_cnt2event_fused: max_count must be in 1..%d", XF_MAXN);
    const int Tw = (int)((S + XF_WSLOTS - 1) / XF_WSLOTS);
    size_t o_hn, o_pre, o_tot, o_key, o_done, o_meta;
    const size_t need = xf_ws_layout(B, Tw, &o_hn, &o_pre, &o_tot, &o_key, &o_done, &o_meta);
    if (workspace_bytes < need) { set_error("esr_cnt2event_fused: workspace %zu < %zu", workspace_bytes, need); return ESR_EWORKSPACE; }
    XfTables tab;
    for (int i = 0; i < XF_NTAB; ++i) {
        const int32_t *d = tables_desc_host + 3 * i;
        ESR_REQUIRE(d[0] >= 0 && d[1] >= 0 && d[2] >= 1 && d[2] <= XF_MAXK && d[0] % 2 == 0 && d[1] % 4 == 0,
                    "esr_cnt2event_fused: bad table descriptor %d", i);
        tab.rank[i] = reinterpret_cast<const uint16_t *>((const char *)tables + d[0]);
        tab.uniq[i] = reinterpret_cast<const float *>((const char *)tables + d[1]);
        tab.K[i] = d[2];
    }
    int ti_cap = 0;
    while ((1 << ti_cap) < max_count) ++ti_cap;
    const int kcap = tab.K[ti_cap];
    const size_t smem = xf_emit_smem(kcap);
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)workspace;
    uint16_t *hn16 = (uint16_t *)(ws + o_hn);
    uint32_t *pre = (uint32_t *)(ws + o_pre), *tot = (uint32_t *)(ws + o_tot), *keybase = (uint32_t *)(ws + o_key);
    unsigned int *done = (unsigned int *)(ws + o_done);
    XfMeta *meta = (XfMeta *)(ws + o_meta);
    ESR_CUDA_CHECK(cudaMemsetAsync(stats, 0, sizeof(int64_t) * 4 * (size_t)B, st));
    ESR_CUDA_CHECK(cudaMemsetAsync(ws, 0, o_hn, st));                         // done counters + meta
    const bool vec = S % 4 == 0 && ((uintptr_t)vals & 15) == 0;
    static bool attr_done = false;
    if (!attr_done) {
        const int most = (int)xf_emit_smem(XF_MAXK);
        ESR_CUDA_CHECK(cudaFuncSetAttribute(k_xf_emit<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, most));
        ESR_CUDA_CHECK(cudaFuncSetAttribute(k_xf_emit<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, most));
        attr_done = true;
    }
    dim3 grid((unsigned)((Tw + XF_WARPS - 1) / XF_WARPS), (unsigned)B);
    unsigned long long *ust = reinterpret_cast<unsigned long long *>(stats);
    const unsigned long long *cst = ust;
    const uint16_t *chn16 = hn16;
    if (vec) ESR_CUDA_CHECK(launch_pdl(k_xf_count<true>, grid, dim3(XF_WARPS * 32), 0, st, vals, S, Tw, hn16, ust));
    else ESR_CUDA_CHECK(launch_pdl(k_xf_count<false>, grid, dim3(XF_WARPS * 32), 0, st, vals, S, Tw, hn16, ust));
    count_launch();
    ESR_CUDA_CHECK(launch_pdl(k_xf_scan, dim3((unsigned)(1 << ti_cap) + 1, (unsigned)B), dim3(256), 0, st, chn16, pre, Tw, B, cst, tab, ti_cap,
                              (unsigned long long)cap_rows, tot, keybase, done, meta));
    count_launch();
    const uint32_t *cpre = pre, *ckey = keybase;
    const XfMeta *cmeta = meta;
    if (vec) ESR_CUDA_CHECK(launch_pdl(k_xf_emit<true>, grid, dim3(XF_WARPS * 32), smem, st, vals, S, Tw, H, W, cpre, ckey, cst, cmeta, tab, kcap, out));
    else ESR_CUDA_CHECK(launch_pdl(k_xf_emit<false>, grid, dim3(XF_WARPS * 32), smem, st, vals, S, Tw, H, W, cpre, ckey, cst, cmeta, tab, kcap, out));
    count_launch();
    return ESR_OK;
}
