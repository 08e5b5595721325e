/* Plain-C caller of libctr_b200.so: no Python, no torch -- the C ABI of include/ctr_b200.h is the whole interface.
 *
 *   c_abi_demo <in.bin> <out.bin>
 *
 * in.bin  : int64 B, F, D, V | int64 field_row_offset[F+1] | int64 ids[B*F] | float table[V*D] | float d_tile[B*F*D] | float d_fm2[B]
 * out.bin : float tile[B*F*D] | float fm2[B] | float row_grads[B*F*D]
 * (tests/test_gpu_c_abi.py writes in.bin, runs this program and checks out.bin against the oracle.)
 *
 * Build: gcc -O2 -I include examples/c_abi_demo.c -o examples/c_abi_demo -L recalgorithm_b200/csrc -lctr_b200 \
 *            -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$ORIGIN/../recalgorithm_b200/csrc' -Wl,-rpath,/usr/local/cuda/lib64
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <cuda_runtime_api.h>

#include "ctr_b200.h"

#define CUDA_OK(call)                                                                      \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 2; } \
  } while (0)
#define CTR_OK_(call)                                                                      \
  do {                                                                                     \
    if ((call) != CTR_OK) { fprintf(stderr, "%s: %s\n", #call, ctr_last_error()); return 3; } \
  } while (0)

static int read_all(FILE* f, void* dst, size_t bytes) { return fread(dst, 1, bytes, f) == bytes ? 0 : -1; }

int main(int argc, char** argv) {
  if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 1; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  int64_t hdr[4];
  if (read_all(f, hdr, sizeof(hdr))) return 1;
  const int64_t B = hdr[0], F = hdr[1], D = hdr[2], V = hdr[3];
  const size_t n_off = (size_t)(F + 1), n_ids = (size_t)(B * F), n_tab = (size_t)(V * D), n_tile = (size_t)(B * F * D);
  int64_t* h_off = malloc(n_off * 8); int64_t* h_ids = malloc(n_ids * 8);
  float* h_tab = malloc(n_tab * 4); float* h_dt = malloc(n_tile * 4); float* h_dg = malloc((size_t)B * 4);
  if (read_all(f, h_off, n_off * 8) || read_all(f, h_ids, n_ids * 8) || read_all(f, h_tab, n_tab * 4) ||
      read_all(f, h_dt, n_tile * 4) || read_all(f, h_dg, (size_t)B * 4)) { fprintf(stderr, "short input\n"); return 1; }
  fclose(f);

  int sms = 0, major = 0, minor = 0;
  CTR_OK_(ctr_device_info(&sms, &major, &minor));
  fprintf(stderr, "libctr_b200 ABI v%d on sm_%d%d (%d SMs)\n", ctr_version(), major, minor, sms);

  int64_t *d_off, *d_ids; float *d_tab, *d_tile, *d_fm2, *d_dt, *d_dg, *d_rg;
  CUDA_OK(cudaMalloc((void**)&d_off, n_off * 8)); CUDA_OK(cudaMalloc((void**)&d_ids, n_ids * 8));
  CUDA_OK(cudaMalloc((void**)&d_tab, n_tab * 4)); CUDA_OK(cudaMalloc((void**)&d_tile, n_tile * 4));
  CUDA_OK(cudaMalloc((void**)&d_fm2, (size_t)B * 4)); CUDA_OK(cudaMalloc((void**)&d_dt, n_tile * 4));
  CUDA_OK(cudaMalloc((void**)&d_dg, (size_t)B * 4)); CUDA_OK(cudaMalloc((void**)&d_rg, n_tile * 4));
  cudaStream_t st;
  CUDA_OK(cudaStreamCreate(&st));
  CUDA_OK(cudaMemcpyAsync(d_off, h_off, n_off * 8, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(d_ids, h_ids, n_ids * 8, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(d_tab, h_tab, n_tab * 4, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(d_dt, h_dt, n_tile * 4, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(d_dg, h_dg, (size_t)B * 4, cudaMemcpyHostToDevice, st));

  /* fc.input_layer x F + the FM second-order block (DeepFM/deepfm.py:184-200), then the lookup's IndexedSlices gradient */
  CTR_OK_(ctr_embed_fm2_fwd(d_tab, d_off, d_ids, B, F, D, d_tile, d_fm2, st));
  CTR_OK_(ctr_embed_fm2_bwd(d_tile, d_dt, d_dg, B, F, D, d_rg, st));
  /* argument errors come back as codes + text, never as a crash */
  if (ctr_embed_fm2_fwd(d_tab, d_off, d_ids, B, F, 5, d_tile, d_fm2, st) != CTR_ERR_UNSUPPORTED) { fprintf(stderr, "D=5 must be refused\n"); return 4; }
  if (ctr_embed_fm2_fwd(NULL, d_off, d_ids, B, F, D, d_tile, d_fm2, st) != CTR_ERR_INVALID_ARG) { fprintf(stderr, "NULL table must be refused\n"); return 4; }

  float* h_tile = malloc(n_tile * 4); float* h_fm2 = malloc((size_t)B * 4); float* h_rg = malloc(n_tile * 4);
  CUDA_OK(cudaMemcpyAsync(h_tile, d_tile, n_tile * 4, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaMemcpyAsync(h_fm2, d_fm2, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaMemcpyAsync(h_rg, d_rg, n_tile * 4, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  FILE* o = fopen(argv[2], "wb");
  if (!o) { perror(argv[2]); return 1; }
  fwrite(h_tile, 4, n_tile, o); fwrite(h_fm2, 4, (size_t)B, o); fwrite(h_rg, 4, n_tile, o);
  fclose(o);
  fprintf(stderr, "ok: %lld kernels launched\n", (long long)ctr_kernel_launches());
  return 0;
}
