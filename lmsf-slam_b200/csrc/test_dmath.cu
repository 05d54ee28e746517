// test_dmath.cu — device-vs-host check of dmath.cuh: the same IEEE operation sequence must give the same
// bits on the B200 as on the host (the library is built with -fmad=false for exactly this reason).
// Built by the Makefile as csrc/test_dmath, run by tests/test_gpu_parity.py::test_device_algebra_bit_exact.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "dmath.cuh"
#include "fmath.cuh"

using namespace lm;

struct Case {
  double H[36], g[6], A[15], S[9], se3[6], x7[7];
};
struct Out {
  double qr6[6], qr53[3], w6[6], V6[36], w3[3], V3[9], inv[36], chol[6], plus[7];
  int ok_inv, ok_chol;
};

HD void run_case(const Case& c, Out& o) {
  cpqr_solve<6, 6>(c.H, c.g, o.qr6);
  double b[5] = {-1, -1, -1, -1, -1};
  cpqr_solve<5, 3>(c.A, b, o.qr53);
  jacobi_eig<6>(c.H, o.w6, o.V6);
  jacobi_eig<3>(c.S, o.w3, o.V3);
  o.ok_inv = invert6(c.H, o.inv) ? 1 : 0;
  o.ok_chol = spd_solve6(c.H, c.g, o.chol) ? 1 : 0;
  se3_plus(c.x7, c.se3, o.plus);
}

__global__ void k_run(const Case* c, Out* o, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) run_case(c[i], o[i]);
}

// fmath.cuh on the device against the host C library's atan2f / sqrtf (the std::atan2(float, float) / std::sqrt(float)
// the reference's bad-point test calls): bit for bit
__global__ void k_fm(const float2* in, float2* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_float2(atan2f_fdlibm(in[i].x, in[i].y), __fsqrt_rn(fabsf(in[i].x)));
}

static int check_fmath() {
  const int n = 1 << 22;
  float2* h = (float2*)malloc(n * sizeof(float2));
  float2* r = (float2*)malloc(n * sizeof(float2));
  for (int i = 0; i < n; ++i) {
    float sc = (i & 1) ? 100.f : 4.f;
    h[i] = make_float2((rand() / (float)RAND_MAX - 0.5f) * 2 * sc, (rand() / (float)RAND_MAX - 0.5f) * 2 * sc);
  }
  float2 *d_in, *d_out;
  cudaMalloc(&d_in, n * sizeof(float2));
  cudaMalloc(&d_out, n * sizeof(float2));
  cudaMemcpy(d_in, h, n * sizeof(float2), cudaMemcpyHostToDevice);
  k_fm<<<(n + 255) / 256, 256>>>(d_in, d_out, n);
  if (cudaMemcpy(r, d_out, n * sizeof(float2), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  int bad = 0;
  for (int i = 0; i < n; ++i) {
    float a = atan2f(h[i].x, h[i].y), s = sqrtf(fabsf(h[i].x));
    if (memcmp(&a, &r[i].x, 4) != 0 || memcmp(&s, &r[i].y, 4) != 0) {
      if (bad < 5) printf("atan2f(%a, %a): host %a device %a; sqrtf host %a device %a\n", h[i].x, h[i].y, a, r[i].x, s, r[i].y);
      ++bad;
    }
  }
  cudaFree(d_in);
  cudaFree(d_out);
  free(h);
  free(r);
  return bad;
}

static double rnd() { return rand() / (double)RAND_MAX - 0.5; }

int main() {
  const int n = 2000;
  Case* hc = (Case*)malloc(n * sizeof(Case));
  Out* ho = (Out*)malloc(n * sizeof(Out));
  Out* hd = (Out*)malloc(n * sizeof(Out));
  srand(7);
  for (int t = 0; t < n; ++t) {
    Case& c = hc[t];
    double J[40][6];
    int rows = (t % 9 == 0) ? 4 : 40;  // rank-deficient normal matrices too
    memset(c.H, 0, sizeof c.H);
    for (int i = 0; i < rows; ++i)
      for (int j = 0; j < 6; ++j) J[i][j] = rnd() * (j < 3 ? 20 : 1);
    for (int i = 0; i < 6; ++i) {
      c.g[i] = rnd();
      for (int j = 0; j < 6; ++j) {
        double s = 0;
        for (int k = 0; k < rows; ++k) s += J[k][i] * J[k][j];
        c.H[i * 6 + j] = s;
      }
    }
    for (int i = 0; i < 15; ++i) c.A[i] = rnd() * 30;
    if (t % 11 == 0)
      for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 3; ++j) c.A[i * 3 + j] = (i + 1) * (j + 1.5);  // collinear neighbours
    double P[5][3], m[3] = {0, 0, 0};
    for (int i = 0; i < 5; ++i)
      for (int j = 0; j < 3; ++j) {
        P[i][j] = rnd() * (j == 0 ? 2 : 0.05);
        m[j] += P[i][j] / 5;
      }
    memset(c.S, 0, sizeof c.S);
    for (int i = 0; i < 5; ++i)
      for (int r = 0; r < 3; ++r)
        for (int q = 0; q < 3; ++q) c.S[r * 3 + q] += (P[i][r] - m[r]) * (P[i][q] - m[q]);
    for (int i = 0; i < 6; ++i) c.se3[i] = rnd() * (t % 5 == 0 ? 1e-12 : 0.05);
    double qn = 0;
    for (int i = 0; i < 4; ++i) {
      c.x7[i] = rnd();
      qn += c.x7[i] * c.x7[i];
    }
    for (int i = 0; i < 4; ++i) c.x7[i] /= sqrt(qn);
    for (int i = 4; i < 7; ++i) c.x7[i] = rnd() * 10;
  }
  for (int t = 0; t < n; ++t) {
    run_case(hc[t], ho[t]);
    jacobi_eig_generic<3>(hc[t].S, ho[t].w3, ho[t].V3);  // host reference of the scalar 3x3 version = the generic one
    double bb[5] = {-1, -1, -1, -1, -1};
    cpqr_solve_generic<5, 3>(hc[t].A, bb, ho[t].qr53);     // likewise for the index-static 5x3 QR
  }
  Case* dc;
  Out* dout;
  if (cudaMalloc(&dc, n * sizeof(Case)) != cudaSuccess) {
    printf("no device\n");
    return 2;
  }
  cudaMalloc(&dout, n * sizeof(Out));
  cudaMemcpy(dc, hc, n * sizeof(Case), cudaMemcpyHostToDevice);
  k_run<<<(n + 63) / 64, 64>>>(dc, dout, n);
  if (cudaMemcpy(hd, dout, n * sizeof(Out), cudaMemcpyDeviceToHost) != cudaSuccess) {
    printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 3;
  }
  // se3_plus uses sin/cos (libm differs by an ulp or two): compare it with a tolerance, the rest bit for bit
  int bad = 0, bad_plus = 0;
  for (int t = 0; t < n; ++t) {
    if (memcmp(&ho[t], &hd[t], offsetof(Out, plus)) != 0 || ho[t].ok_inv != hd[t].ok_inv || ho[t].ok_chol != hd[t].ok_chol) {
      if (bad < 5) {
        const char* names[] = {"qr6", "qr53", "w6", "V6", "w3", "V3", "inv", "chol"};
        size_t offs[] = {offsetof(Out, qr6), offsetof(Out, qr53), offsetof(Out, w6), offsetof(Out, V6), offsetof(Out, w3),
                         offsetof(Out, V3), offsetof(Out, inv), offsetof(Out, chol), offsetof(Out, plus)};
        for (int k = 0; k < 8; ++k)
          if (memcmp((char*)&ho[t] + offs[k], (char*)&hd[t] + offs[k], offs[k + 1] - offs[k]) != 0)
            printf("case %d: %s differs (host %.17g device %.17g)\n", t, names[k], *(double*)((char*)&ho[t] + offs[k]),
                   *(double*)((char*)&hd[t] + offs[k]));
      }
      ++bad;
    }
    for (int i = 0; i < 7; ++i)
      if (fabs(ho[t].plus[i] - hd[t].plus[i]) > 1e-14 * (1 + fabs(ho[t].plus[i]))) ++bad_plus;
  }
  int bad_fm = check_fmath();
  printf("cases %d  bit mismatches %d  se3_plus out of tolerance %d  atan2f/sqrtf mismatches (4M arguments) %d\n", n, bad, bad_plus,
         bad_fm);
  return (bad || bad_plus || bad_fm) ? 1 : 0;
}
