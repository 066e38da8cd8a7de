// Bias probe: mean and rms of c^2 + |w|^2 - 1 over random rotations for several formulations.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "jacobi_wide.cuh"
using namespace nls;
__global__ void rot_kernel(const double* in, double* out, int n, int mode) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a = in[4 * i], b = in[4 * i + 1], zr = in[4 * i + 2], zi = in[4 * i + 3];
  const double az2 = zr * zr + zi * zi;
  double c, wr, wi;
  if (mode == 0) {
    jacobi_rotation(a, b, zr, zi, az2, c, wr, wi);
  } else if (mode == 1) {
    const double inv_az = rsqrt(az2);
    const double tau = 0.5 * (b - a) * inv_az;
    const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
    c = rsqrt(1.0 + t * t);
    const double s = t * c;
    wr = s * zr * inv_az;
    wi = s * zi * inv_az;
  }
  out[3 * i] = c; out[3 * i + 1] = wr; out[3 * i + 2] = wi;
}
__global__ void rs_kernel(const double* in, double* out, int n, int mode) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = fabs(in[i]) + 1.0;
  out[2 * i] = x;
  out[2 * i + 1] = mode == 0 ? rsqrt_seeded(x) : rsqrt(x);
}
int main() {
  const int n = 1 << 20;
  std::vector<double> in(4 * n), out(3 * n);
  srand(2);
  for (int i = 0; i < n; ++i) {
    in[4 * i] = (rand() / (double)RAND_MAX - 0.5) * 8; in[4 * i + 1] = (rand() / (double)RAND_MAX - 0.5) * 8;
    const double sc = pow(10.0, -6.0 * rand() / (double)RAND_MAX);
    in[4 * i + 2] = (rand() / (double)RAND_MAX - 0.5) * sc; in[4 * i + 3] = (rand() / (double)RAND_MAX - 0.5) * sc;
  }
  double *di, *dout;
  cudaMalloc(&di, in.size() * 8); cudaMalloc(&dout, out.size() * 8);
  cudaMemcpy(di, in.data(), in.size() * 8, cudaMemcpyHostToDevice);
  const char* names[] = {"jacobi_rotation", "textbook"};
  for (int mode = 0; mode < 2; ++mode) {
    rot_kernel<<<n / 256, 256>>>(di, dout, n, mode);
    cudaMemcpy(out.data(), dout, out.size() * 8, cudaMemcpyDeviceToHost);
    long double sum = 0, sq = 0, mx = 0;
    for (int i = 0; i < n; ++i) {
      const long double c = out[3 * i], wr = out[3 * i + 1], wi = out[3 * i + 2];
      const long double e = c * c + wr * wr + wi * wi - 1.0L;
      sum += e; sq += e * e; mx = fmaxl(mx, fabsl(e));
    }
    printf("%-24s: mean(n2-1) %+.3Le  rms %.3Le  max %.3Le\n", names[mode], sum / n, sqrtl(sq / n), mx);
  }
  for (int mode = 0; mode < 2; ++mode) {
    rs_kernel<<<n / 256, 256>>>(di, dout, n, mode);
    cudaMemcpy(out.data(), dout, 2 * n * 8, cudaMemcpyDeviceToHost);
    long double sum = 0, sq = 0;
    for (int i = 0; i < n; ++i) {
      const long double x = out[2 * i], y = out[2 * i + 1];
      const long double e = y * sqrtl(x) - 1.0L;
      sum += e; sq += e * e;
    }
    printf("rsqrt %s: mean rel err %+.3Le rms %.3Le\n", mode == 0 ? "seeded" : "lib   ", sum / n, sqrtl(sq / n));
  }
  return 0;
}
