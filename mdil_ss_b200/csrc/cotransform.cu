// GPU side of the training co-transform (reference train_new_task_step2.py:48-81 MyCoTransform,
// transform.py:63-79 Relabel / ToLabel): uint8 source image + label -> resized (Pillow BILINEAR fixed-point two-pass
// resampling for the image, NEAREST for the label), horizontally flipped, translated by -2..2 pixels with the reference's
// fill rules, float [0,1] NCHW image + int64 label with 255 relabelled to C-1.  Bit-exact with the reference (Pillow's
// 22-bit coefficients and its 8-bit intermediate between the two passes are reproduced): one thread per output pixel.
#include "kernels.cuh"

namespace mdil {

namespace {
constexpr int kPrecisionBits = 32 - 8 - 2;   // Pillow Resample.c PRECISION_BITS

__device__ __forceinline__ int clip8(int v) {
  v >>= kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// tab: [out][2 + K] = first source index, tap count, K coefficients
__global__ void __launch_bounds__(256)
cotransform_kernel(const unsigned char* __restrict__ img, const unsigned char* __restrict__ lab, int N, int Hs, int Ws, int H,
                   int W, const int* __restrict__ xtab, int KX, const int* __restrict__ ytab, int KY,
                   const int* __restrict__ xnear, const int* __restrict__ ynear, const int* __restrict__ params,
                   int num_classes, float* __restrict__ out_img, long long* __restrict__ out_lab) {
  const long total = (long)N * H * W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H), n = (int)(i / ((long)W * H));
    int flip = 0, tx = 0, ty = 0;
    if (params != nullptr) { flip = params[3 * n]; tx = params[3 * n + 1]; ty = params[3 * n + 2]; }
    int r = 0, g = 0, b = 0, l = 0;
    // ImageOps.expand(border=(tx, ty, 0, 0)) then crop((0, 0, W, H)): beyond the expanded image the crop pads with 0
    // (image and label), the expand border is filled with 0 (image) / 255 (label)
    if (x < W + tx && y < H + ty) {
      const int sx = x - tx, sy = y - ty;
      if (sx < 0 || sy < 0) {
        l = 255;
      } else {
        const int rx = flip ? W - 1 - sx : sx, ry = sy;
        const int* xt = xtab + (long)rx * (2 + KX);
        const int* yt = ytab + (long)ry * (2 + KY);
        const int x0 = xt[0], nx = xt[1], y0 = yt[0], ny = yt[1];
        int ar = 1 << (kPrecisionBits - 1), ag = ar, ab = ar;
        const unsigned char* base = img + (long)n * Hs * Ws * 3;
        for (int j = 0; j < ny; ++j) {
          const unsigned char* row = base + ((long)(y0 + j) * Ws + x0) * 3;
          int hr = 1 << (kPrecisionBits - 1), hg = hr, hb = hr;
          for (int k = 0; k < nx; ++k) {
            const int c = xt[2 + k];
            hr += (int)row[3 * k] * c; hg += (int)row[3 * k + 1] * c; hb += (int)row[3 * k + 2] * c;
          }
          const int cy = yt[2 + j];
          ar += clip8(hr) * cy; ag += clip8(hg) * cy; ab += clip8(hb) * cy;
        }
        r = clip8(ar); g = clip8(ag); b = clip8(ab);
        l = lab[((long)n * Hs + ynear[ry]) * Ws + xnear[rx]];
      }
    }
    const long plane = (long)H * W, o = (long)n * 3 * plane + (long)y * W + x;
    out_img[o] = (float)r / 255.0f;
    out_img[o + plane] = (float)g / 255.0f;
    out_img[o + 2 * plane] = (float)b / 255.0f;
    out_lab[i] = l == 255 ? num_classes - 1 : l;
  }
}
}  // namespace

int launch_cotransform(const unsigned char* img, const unsigned char* lab, int N, int Hs, int Ws, int H, int W, const int* xtab,
                       int KX, const int* ytab, int KY, const int* xnear, const int* ynear, const int* params, int num_classes,
                       float* out_img, long long* out_lab, cudaStream_t s) {
  MDIL_REQUIRE(N > 0 && Hs > 0 && Ws > 0 && H > 0 && W > 0 && KX > 0 && KY > 0 && num_classes > 0 && num_classes <= 255,
               "cotransform: bad arguments");
  const long total = (long)N * H * W;
  long grid = (total + 255) / 256;
  if (grid > kNumSMs * 16) grid = kNumSMs * 16;
  cotransform_kernel<<<(int)grid, 256, 0, s>>>(img, lab, N, Hs, Ws, H, W, xtab, KX, ytab, KY, xnear, ynear, params, num_classes,
                                               out_img, out_lab);
  MDIL_LAUNCH_CHECK();
  return 0;
}

}  // namespace mdil
