// TMA load throughput into ONE SM as a function of the box shape, for the activation-plane layout of the conv kernels
// ([images][8 chunks][H][W][8 ch] fp16: a pixel's 8 channels are 16 B, an image row of a chunk is W*16 B).
// Every box moves 16 KB (128 pixels x 64 channels); what changes is how long its contiguous rows are:
//   8 px  x 16 rows x 8 chunks = 128 rows of 128 B   (the 16x8 tile of the 1x1 conv10: pfrb_flow.cu)
//   16 px x  8 rows x 8 chunks =  64 rows of 256 B
//   32 px x  4 rows x 8 chunks =  32 rows of 512 B
//   10 px x 18 rows x 8 chunks = 144 rows of 160 B   (the halo patch of the 3x3 convs, 22.5 KB)
// conv10 tiles take 7.0 K cycles for 14 such boxes (32 B/clk) although their MMAs need 4.1 K and a 7th ring slot
// changed nothing (profiles/r2z_flow_balance.txt): is that the rate at which the TMA unit emits 128-byte rows?
// One producer thread per CTA keeps `depth` boxes in flight (ring of mbarriers), nobody reads the data.
// The source (16 images of 64x64, 8 MB) stays in the L2.  Grid: 148 CTAs (all SMs pulling) or 10 (conv10's share).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o probes/bin/tma_rate_probe probes/tma_rate_probe.cu
#include <stdio.h>
#include <stdlib.h>

#include "../pfnl_b200/csrc/tc_ptx.cuh"
#include "../pfnl_b200/csrc/tc_tmap.h"

using namespace pfnl::tc;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

constexpr int kSlots = 8;
constexpr int kSlotBytes = 23552;

__global__ void __launch_bounds__(64, 1)
    tma_kernel(const __grid_constant__ CUtensorMap tm, int box_w, int box_h, int box_bytes, int n_boxes, int depth, int H,
               int W, int images, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[kSlots];
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int tiles_x = W / box_w, tiles_y = H / box_h;
  // no divisions inside the timed loop (a first version measured its own index arithmetic: 772 cycles per box)
  int tx = blockIdx.x % tiles_x, ty = (blockIdx.x / tiles_x) % tiles_y, img = blockIdx.x % images;
  int slot_i = 0, par_i = 0;  // slot / parity of the box being issued
  int slot_w = 0, par_w = 0;  // ... of the box being waited for
  const long long t0 = clock64();
  bool ok = true;
  for (int i = 0; i < n_boxes + depth && ok; ++i) {
    if (i >= depth) {  // box i - depth has to land before its slot is reused (bounded: a mistake must not hang the box)
      bool done = false;
      for (long long spin = 0; spin < (1ll << 22) && !done; ++spin) done = mbar_try_wait(&full[slot_w], par_w);
      ok = done;
      if (++slot_w == depth) slot_w = 0, par_w ^= 1;
    }
    if (i < n_boxes && ok) {
      mbar_arrive_expect_tx(&full[slot_i], box_bytes);
      tma_load_4d(smem + slot_i * kSlotBytes, &tm, &full[slot_i], tx * box_w * 8, ty * box_h, 0, img);
      if (++slot_i == depth) slot_i = 0, par_i ^= 1;
      if (++tx == tiles_x) {
        tx = 0;
        if (++ty == tiles_y) {
          ty = 0;
          if (++img == images) img = 0;
        }
      }
    }
  }
  (void)par_i;
  out[blockIdx.x] = ok ? clock64() - t0 : -1;
}

static void run(const char* label, void* src, int images, int H, int W, int box_w, int box_h, int depth, int grid,
                long long* d_out) {
  CUtensorMap tm;
  if (make_act_tmap(&tm, src, images, H, W, box_w, box_h) != 0) {
    printf("%-28s: cuTensorMapEncodeTiled failed\n", label);
    return;
  }
  const int box_bytes = box_w * box_h * 128, n_boxes = 2000;
  const int smem = 1024 + kSlots * kSlotBytes;
  CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int rep = 0; rep < 2; ++rep) {
    tma_kernel<<<grid, 64, smem>>>(tm, box_w, box_h, box_bytes, n_boxes, depth, H, W, images, d_out);
    CK(cudaDeviceSynchronize());
  }
  long long h[256];
  CK(cudaMemcpy(h, d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0;
  int bad = 0;
  for (int i = 0; i < grid; ++i) {
    if (h[i] < 0) ++bad;
    avg += (double)h[i];
  }
  avg /= grid;
  printf("%-28s box %2d px x %2d rows (%3d rows of %3d B, %5d B) depth %d grid %3d: %7.1f cycles/box  %5.1f B/clk/SM%s\n", label,
         box_w, box_h, box_h * 8, box_w * 16, box_bytes, depth, grid, avg / n_boxes, box_bytes / (avg / n_boxes),
         bad ? "  [TIMEOUT]" : "");
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int images = 16, H = 64, W = 64;
  void* src;
  CK(cudaMalloc(&src, (size_t)images * 8 * H * W * 16));
  CK(cudaMemset(src, 0, (size_t)images * 8 * H * W * 16));
  long long* d_out;
  CK(cudaMalloc(&d_out, 256 * sizeof(long long)));
  printf("== tma_rate_probe: %s, %d SMs; source %d x [8][%d][%d][8] fp16 = %.1f MB (L2 resident) ==\n", prop.name,
         prop.multiProcessorCount, images, H, W, images * 8.0 * H * W * 16 / 1e6);
  const int grids[2] = {prop.multiProcessorCount, 10};
  for (int g = 0; g < 2; ++g)
    for (int depth = 2; depth <= 8; depth += 2) {
      run("conv10 tile (16x8)", src, images, H, W, 8, 16, depth, grids[g], d_out);
      run("16 px rows", src, images, H, W, 16, 8, depth, grids[g], d_out);
      run("32 px rows", src, images, H, W, 32, 4, depth, grids[g], d_out);
      run("3x3 halo patch (18x10)", src, images, H, W, 10, 18, depth > 6 ? 6 : depth, grids[g], d_out);
    }
  printf("== done ==\n");
  return 0;
}
