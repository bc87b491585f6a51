// Audio front-end of the path's caller (inference.py:112-113,230-231): torchaudio.transforms.Resample(sr, 16000)(audio)
// followed by .mean(dim=0) -- a polyphase windowed-sinc FIR (torchaudio/functional/functional.py,
// _get_sinc_resample_kernel / _apply_sinc_resample_kernel: zero padding (width, width + orig), conv1d with stride orig,
// output cut to ceil(new * length / orig)) and the channel mix, as ONE pass over the input on the device:
//   out[i * new + p] = 1/C * sum_c sum_k K[p][k] * x_c[i * orig + k - width]          (zero outside [0, length))
// HBM-bound (4 B read per input sample per channel, 4 B written per output sample); the filter bank K (new x taps floats,
// built on the host in fp64 like torchaudio does) is read through the read-only cache. One thread per output sample; a
// warp's 32 consecutive outputs read overlapping, contiguous input windows.
#include "kernels.cuh"

namespace artalk {

__global__ void __launch_bounds__(256) resample_mix_kernel(const float* __restrict__ in, int channels, int64_t ch_stride, int64_t length,
                                                           const float* __restrict__ bank, int orig, int new_f, int taps, int width,
                                                           float* __restrict__ out, int64_t out_len) {
  pdl_enter();
  const float inv_c = 1.0f / (float)channels;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < out_len; n += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = n / new_f;
    const int ph = (int)(n - i * new_f);
    const float* k = bank + (int64_t)ph * taps;
    const int64_t s0 = i * orig - width;                    // first input sample under the filter
    int k_lo = 0, k_hi = taps;
    if (s0 < 0) k_lo = (int)(-s0);
    if (s0 + taps > length) k_hi = (int)(length - s0);
    float mix = 0.f;
    for (int c = 0; c < channels; ++c) {
      const float* x = in + c * ch_stride + s0;
      float acc = 0.f;
      for (int j = k_lo; j < k_hi; ++j) acc = fmaf(__ldg(k + j), x[j], acc);
      mix += acc;
    }
    out[n] = mix * inv_c;
  }
}

int launch_resample_mix(const float* in, int channels, int64_t ch_stride, int64_t length, const float* bank, int orig, int new_f,
                        int taps, int width, float* out, int64_t out_len, cudaStream_t st) {
  if (out_len <= 0) return AT_OK;
  AT_REQUIRE(in && bank && out && channels >= 1 && length >= 1, "resample: bad argument");
  AT_REQUIRE(orig >= 1 && new_f >= 1 && taps == 2 * width + orig && width >= 1, "resample: inconsistent filter bank (taps = 2 * width + orig)");
  AT_REQUIRE(out_len <= (length * new_f + orig - 1) / orig, "resample: out_len exceeds ceil(new * length / orig)");
  int64_t blocks = (out_len + 255) / 256;
  int grid = (int)(blocks < 148 * 16 ? blocks : 148 * 16);
  AT_CUDA(launch_k(resample_mix_kernel, dim3(grid), dim3(256), 0, st, in, channels, ch_stride, length, bank, orig, new_f, taps, width,
                   out, out_len));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

}  // namespace artalk
