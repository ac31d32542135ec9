// reader.cu -- device side of the columnar event reader: window indexing by timestamp search and the
// gather of per-frame event slices from the (pinned, host-mapped or device-resident) columns straight into the fp32 SoA the
// count-scatter kernels consume.
//
// Replaces, for whole batches at once:
//   BaseDataset.binary_search_h5_dset   dataloader/base_dataset.py:78-91 (= dataloader/binary_search/binary_search.pyx:17-38):
//       bisection over the sorted float64 timestamps that returns `mid` as soon as dset[mid] == x, else the left insertion
//       point -- with duplicate timestamps this is NOT numpy.searchsorted; the same probe sequence is reproduced so the indices
//       are bit-identical;
//   H5Dataset.get_events / get_gt_events + BaseDataset.event_formatting   dataloader/h5dataset.py:492-506,
//       dataloader/base_dataset.py:26-33: int16 x / y and float64 t / p slices -> float32, t normalised per frame
//       (ts - ts[0]) / (ts[-1] - ts[0] + 1e-6) in float32 arithmetic, as torch evaluates it.
#include "common.cuh"

namespace esr {

__global__ void __launch_bounds__(256)
k_ts_search(const double *__restrict__ ts, long long n, const double *__restrict__ q, long long nq, long long *__restrict__ out)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (long long)gridDim.x * blockDim.x) {
        const double x = q[i];
        long long l = 0, r = n - 1, res = -1;
        while (l <= r) {
            const long long mid = l + (r - l) / 2;
            const double v = ts[mid];
            if (v == x) { res = mid; break; }
            if (v < x) l = mid + 1; else r = mid - 1;
        }
        out[i] = res >= 0 ? res : l;
    }
}

// one block per frame: events [start[f], start[f] + len[f]) of the columns -> out[off[f] ...] as fp32
__global__ void __launch_bounds__(256)
k_gather_events(const short *__restrict__ xs, const short *__restrict__ ys, const double *__restrict__ ts,
                const double *__restrict__ ps, const long long *__restrict__ start, const long long *__restrict__ off,
                float *__restrict__ oxs, float *__restrict__ oys, float *__restrict__ ots, float *__restrict__ ops)
{
    const int f = blockIdx.x;
    const long long s = start[f], o = off[f], n = off[f + 1] - o;
    float t0 = 0.f, den = 1.f;
    if (ots && n > 0) {
        t0 = (float)ts[s];
        den = __fadd_rn(__fsub_rn((float)ts[s + n - 1], t0), 1e-6f);          // fp32, like the torch expression
    }
    for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.y * blockDim.x) {
        oxs[o + i] = (float)xs[s + i];
        oys[o + i] = (float)ys[s + i];
        ops[o + i] = (float)ps[s + i];
        if (ots) ots[o + i] = __fdiv_rn(__fsub_rn((float)ts[s + i], t0), den);
    }
}

} // namespace esr

using namespace esr;

extern "C" int esr_ts_search(const double *ts, int64_t n, const double *queries, int64_t nq, int64_t *out, esr_stream_t stream)
{
    ESR_REQUIRE(ts && queries && out && n >= 0 && nq >= 0, "esr_ts_search: bad arguments");
    if (nq == 0) return ESR_OK;
    k_ts_search<<<(unsigned)min((int64_t)4096, (nq + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ts, (long long)n, queries, (long long)nq,
                                                                                              (long long *)out);
    ESR_LAUNCH_CHECK();
    return ESR_OK;
}

extern "C" int esr_gather_events(const int16_t *xs, const int16_t *ys, const double *ts, const double *ps, const int64_t *start,
                                 const int64_t *off, int n_frames, int64_t max_len, float *out_xs, float *out_ys, float *out_ts,
                                 float *out_ps, esr_stream_t stream)
{
    ESR_REQUIRE(xs && ys && ps && start && off && out_xs && out_ys && out_ps && n_frames >= 0, "esr_gather_events: bad arguments");
    ESR_REQUIRE(!out_ts || ts, "esr_gather_events: out_ts needs the ts column");
    if (n_frames == 0) return ESR_OK;
    ESR_REQUIRE(n_frames <= 0x7fffffff, "esr_gather_events: too many frames");
    const unsigned gy = (unsigned)max((int64_t)1, min((int64_t)64, (max_len + 2047) / 2048));
    k_gather_events<<<dim3((unsigned)n_frames, gy), 256, 0, (cudaStream_t)stream>>>(xs, ys, ts, ps, (const long long *)start, (const long long *)off,
                                                                                out_xs, out_ys, out_ts, out_ps);
    ESR_LAUNCH_CHECK();
    return ESR_OK;
}
