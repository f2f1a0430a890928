/* A plain C host for the C ABI of include/soda_cuda.h: what a caller of the
 * reference's soda::app::jacobi2d(ptr, extent, stride, min, ..., bitstream, ...)
 * (src/soda/codegen/frt/host.py:62-89) writes against this backend instead.
 *
 *   jacobi2d_host WIDTH HEIGHT input.bin output.bin
 *
 * input.bin / output.bin: HEIGHT x WIDTH float32, dimension 0 contiguous.  The
 * output file is pre-filled with the caller's values; only the valid interior
 * is overwritten (src/soda/codegen/frt/host.py:357-374). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "soda_cuda.h"

int soda_cuda_jacobi2d(const float* var_t1_ptr, const int32_t* var_t1_extent,
                       const int32_t* var_t1_stride, const int32_t* var_t1_min,
                       float* var_t0_ptr, const int32_t* var_t0_extent,
                       const int32_t* var_t0_stride, const int32_t* var_t0_min,
                       const soda_cuda_opts* opts);

static float* read_all(const char* path, size_t count) {
  FILE* fp = fopen(path, "rb");
  float* data = (float*)malloc(count * sizeof(float));
  if (fp == NULL || data == NULL || fread(data, sizeof(float), count, fp) != count) {
    fprintf(stderr, "cannot read %s\n", path);
    exit(2);
  }
  fclose(fp);
  return data;
}

int main(int argc, char** argv) {
  if (argc != 5) {
    fprintf(stderr, "usage: %s WIDTH HEIGHT input.bin output.bin\n", argv[0]);
    return 2;
  }
  const int32_t width = atoi(argv[1]), height = atoi(argv[2]);
  const size_t count = (size_t)width * height;
  float* in = read_all(argv[3], count);
  float* out = read_all(argv[4], count);

  soda_cuda_program_info info;
  if (soda_cuda_info(&info) != SODA_CUDA_OK) return 3;
  printf("program %s: %d-D, iterate %d, %d pass(es)\n", info.app_name, info.dim,
         info.iterate, info.num_passes);

  int32_t extent[2] = {width, height}, stride[2] = {1, width}, min[2] = {0, 0};
  int status = soda_cuda_jacobi2d(in, extent, stride, min, out, extent, stride,
                                  min, NULL);
  if (status != SODA_CUDA_OK) {
    fprintf(stderr, "soda_cuda_jacobi2d: status %d: %s\n", status,
            soda_cuda_last_error());
    return 1;
  }
  FILE* fp = fopen(argv[4], "wb");
  if (fp == NULL || fwrite(out, sizeof(float), count, fp) != count) return 2;
  fclose(fp);
  printf("launches: %lld\n", (long long)soda_cuda_launch_count());
  free(in);
  free(out);
  return 0;
}
