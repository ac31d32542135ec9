// metrics.cu -- evaluation metrics of infer_ours_cnt.py:81-100 on the GPU: per (sample, channel) plane
//   sum |pred - tgt|, sum (pred - tgt)^2, max / min of tgt        -> L1, MSE, PSNR (loss/restore.py:64-90)
//   sum over the valid region of the SSIM map, 7x7 uniform window  -> SSIM (loss/restore.py:42-61 -> skimage
//   structural_similarity: sample covariance NP/(NP-1), K1 = 0.01, K2 = 0.03, window means over a (win x win) box, mean of
//   S over the image cropped by (win-1)/2)
// The reference moves both tensors to the CPU and calls skimage per channel; here one launch pair serves every plane of a
// batch.  Window sums and reductions run in fp64 (uxx - ux^2 cancels ~5 digits at count-image magnitudes; skimage itself
// computes in float64), reductions are two-stage and deterministic (no atomics).
#include "common.cuh"

namespace esr {

constexpr int MT = 16;                       // SSIM output tile (MT x MT pixels per block)
constexpr int M_MAXWIN = 11;

// block-wide sum / max / min of doubles (256 threads)
__device__ __forceinline__ double block_sum(double v, double *sh)
{
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    const double r = sh[0];
    __syncthreads();
    return r;
}

// one block per plane: {sum |d|, sum d^2, max tgt, min tgt}
__global__ void __launch_bounds__(256)
k_metric_plane_stats(const float *__restrict__ pred, const float *__restrict__ tgt, int HW, double *__restrict__ out /*[planes][6]*/)
{
    __shared__ double sh[256];
    const size_t base = (size_t)blockIdx.x * HW;
    double sa = 0.0, sq = 0.0, mx = -1e300, mn = 1e300;
    for (int i = threadIdx.x; i < HW; i += 256) {
        const double p = pred[base + i], t = tgt[base + i], d = p - t;
        sa += fabs(d); sq += d * d; mx = fmax(mx, t); mn = fmin(mn, t);
    }
    sa = block_sum(sa, sh); sq = block_sum(sq, sh);
    sh[threadIdx.x] = mx; __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + o]); __syncthreads(); }
    mx = sh[0]; __syncthreads();
    sh[threadIdx.x] = mn; __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] = fmin(sh[threadIdx.x], sh[threadIdx.x + o]); __syncthreads(); }
    mn = sh[0];
    if (threadIdx.x == 0) { double *o = out + (size_t)blockIdx.x * 6; o[0] = sa; o[1] = sq; o[2] = mx; o[3] = mn; }
}

// SSIM map over the valid region [pad, H-pad) x [pad, W-pad): one block per (tile, plane) -> partial sum of S
__global__ void __launch_bounds__(256)
k_ssim_tiles(const float *__restrict__ pred, const float *__restrict__ tgt, int H, int W, int win, double C1, double C2,
             int tiles_x, double *__restrict__ partial /*[planes][tiles]*/)
{
    __shared__ double sx[(MT + M_MAXWIN - 1) * (MT + M_MAXWIN - 1)], sy[(MT + M_MAXWIN - 1) * (MT + M_MAXWIN - 1)];
    __shared__ double sh[256];
    const int pad = (win - 1) / 2, ext = MT + win - 1;
    const int tile = blockIdx.x, plane = blockIdx.y;
    const int ty0 = pad + (tile / tiles_x) * MT, tx0 = pad + (tile % tiles_x) * MT;      // first valid output pixel of the tile
    const float *P = pred + (size_t)plane * H * W, *T = tgt + (size_t)plane * H * W;
    for (int i = threadIdx.x; i < ext * ext; i += 256) {
        const int y = ty0 - pad + i / ext, x = tx0 - pad + i % ext;
        const bool in = y < H && x < W;                                               // y, x >= 0 by construction
        sx[i] = in ? (double)P[(size_t)y * W + x] : 0.0;
        sy[i] = in ? (double)T[(size_t)y * W + x] : 0.0;
    }
    __syncthreads();
    const int ly = threadIdx.x / MT, lx = threadIdx.x % MT;
    const int y = ty0 + ly, x = tx0 + lx;
    double S = 0.0;
    if (y < H - pad && x < W - pad) {
        double ux = 0, uy = 0, uxx = 0, uyy = 0, uxy = 0;
        for (int i = 0; i < win; ++i)
            for (int j = 0; j < win; ++j) {
                const double a = sx[(ly + i) * ext + lx + j], b = sy[(ly + i) * ext + lx + j];
                ux += a; uy += b; uxx += a * a; uyy += b * b; uxy += a * b;
            }
        const double NP = (double)(win * win), inv = 1.0 / NP, cov = NP / (NP - 1.0);
        ux *= inv; uy *= inv; uxx *= inv; uyy *= inv; uxy *= inv;
        const double vx = cov * (uxx - ux * ux), vy = cov * (uyy - uy * uy), vxy = cov * (uxy - ux * uy);
        S = ((2.0 * ux * uy + C1) * (2.0 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
    }
    const double tot = block_sum(S, sh);
    if (threadIdx.x == 0) partial[(size_t)plane * gridDim.x + tile] = tot;
}

__global__ void __launch_bounds__(256)
k_ssim_finish(const double *__restrict__ partial, int tiles, double count, double *__restrict__ out /*[planes][6]*/)
{
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < tiles; i += 256) s += partial[(size_t)blockIdx.x * tiles + i];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) { out[(size_t)blockIdx.x * 6 + 4] = s; out[(size_t)blockIdx.x * 6 + 5] = count; }
}

} // namespace esr

using namespace esr;

static void ssim_geo(int H, int W, int win, int *tiles_x, int *tiles)
{
    const int pad = (win - 1) / 2;
    const int vh = H - 2 * pad, vw = W - 2 * pad;
    *tiles_x = vw > 0 ? (vw + MT - 1) / MT : 0;
    *tiles = (vh > 0 && vw > 0) ? *tiles_x * ((vh + MT - 1) / MT) : 0;
}

extern "C" size_t esr_metrics_workspace_bytes(int n_planes, int H, int W, int win)
{
    int tx, t;
    ssim_geo(H, W, win, &tx, &t);
    return align_up((size_t)n_planes * (size_t)(t > 0 ? t : 1) * sizeof(double), 256);
}

extern "C" int esr_metrics_planes(const float *pred, const float *tgt, int n_planes, int H, int W, int win, double data_range,
                                  double *stats, void *workspace, size_t ws_bytes, esr_stream_t stream)
{
    ESR_REQUIRE(pred && tgt && stats && workspace, "esr_metrics_planes: null pointer");
    ESR_REQUIRE(n_planes > 0 && H > 0 && W > 0 && (long long)H * W < (1ll << 31), "esr_metrics_planes: bad dims");
    ESR_REQUIRE(win >= 3 && win <= M_MAXWIN && (win & 1) && win <= H && win <= W,
                "esr_metrics_planes: win_size %d must be odd, in [3, %d] and not exceed the image side (skimage raises ValueError)", win, M_MAXWIN);
    ESR_REQUIRE(n_planes <= 65535, "esr_metrics_planes: too many planes");
    const size_t need = esr_metrics_workspace_bytes(n_planes, H, W, win);
    if (ws_bytes < need) { set_error("esr_metrics_planes: workspace %zu < %zu", ws_bytes, need); return ESR_EWORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    int tiles_x, tiles;
    ssim_geo(H, W, win, &tiles_x, &tiles);
    const int pad = (win - 1) / 2;
    k_metric_plane_stats<<<n_planes, 256, 0, st>>>(pred, tgt, H * W, stats);
    ESR_LAUNCH_CHECK();
    const double C1 = (0.01 * data_range) * (0.01 * data_range), C2 = (0.03 * data_range) * (0.03 * data_range);
    k_ssim_tiles<<<dim3(tiles, n_planes), 256, 0, st>>>(pred, tgt, H, W, win, C1, C2, tiles_x, (double *)workspace);
    ESR_LAUNCH_CHECK();
    k_ssim_finish<<<n_planes, 256, 0, st>>>((const double *)workspace, tiles, (double)(H - 2 * pad) * (double)(W - 2 * pad), stats);
    ESR_LAUNCH_CHECK();
    return ESR_OK;
}
