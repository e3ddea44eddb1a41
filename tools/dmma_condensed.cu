// Measurement, not product code: the CONDENSED Newton system of the MPC on FP64 tensor cores
// (mma.sync.m8n8k4.f64 = DMMA), to be set against the stage-wise Riccati sweep of
// csrc/ipm_quad.cuh (BASELINE.json north_star: "tensor cores used only for the per-step dense KKT
// factor where it is a true small GEMM"; SURVEY.md section 7: "measure before committing").
//
// The dynamics are affine with constant Phi, Gam, so the states are a constant linear map of the
// controls and the Newton system condenses to the controls alone.  The yaw chain decouples
// (20 x 20, diagonal-plus-rank structure); the three axis chains are coupled by the 6x6 (p, v)
// Hessian blocks of the collision terms.  With t_a(l) = F_a^l G_a (3-vector: response of chain a,
// lag l) the reduced Hessian of the axis controls is, for axis pair (a, a'),
//     H^(aa')[j, j'] = sum_{k > max(j, j')}  t_a(k-1-j)' Q_k^(aa') t_a'(k-1-j')       (20 x 20)
//                    = A^(a) W^(aa'),   A^(a) = [A_1 .. A_N] (20 x 3N, row j of A_k = t_a(k-1-j)'),
//                                       W^(aa') = blockdiag(Q_k^(aa')) A^(a')'        (3N x 20)
// i.e. nine 20 x 60 x 20 GEMMs per instance and iteration (3 x 3 tiles of 8 x 8, 15 k-steps of 4
// each), then a 60 x 60 Cholesky.  This program times
//   (1) the nine GEMMs on DMMA, one warp per instance, operands staged in shared memory,
//   (2) a right-looking blocked Cholesky of the 60 x 60 result (8 x 8 diagonal blocks scalar,
//       trailing updates on DMMA),
// at full occupancy on all SMs, and prints cycles per instance per SM for both.  The Riccati
// figure to compare with comes from the ncu profile of the solve kernel (profiles/).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define N_STAGES 20
#define KDIM 60 // 3 N
#define HP 24   // 20 padded to 3 tiles of 8

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// shared memory per warp: A (3 chains x 24 x 60), W (60 x 24, one pair at a time), H (72 x 72)
struct WarpSmem {
    double A[3][HP][KDIM + 4]; // +4: bank spread
    double W[KDIM][HP + 1];
    double H[72][73];
};

__global__ void __launch_bounds__(32) condensed_kernel(const double *__restrict__ tresp /* [3][N][3] */,
                                                       const double *__restrict__ Q /* [inst][N][3][3][3][3] */,
                                                       double *__restrict__ out, int n_inst, int reps, int do_chol) {
    extern __shared__ __align__(16) unsigned char raw[];
    WarpSmem &S = *reinterpret_cast<WarpSmem *>(raw);
    const int lane = threadIdx.x;
    // A^(a): row j of block k (columns 3k..3k+2) = t_a(k - j) for j <= k (k = 0..N-1 here), else 0
    for (int e = lane; e < 3 * HP * KDIM; e += 32) {
        const int a = e / (HP * KDIM), r = (e / KDIM) % HP, col = e % KDIM, k = col / 3, c = col % 3;
        S.A[a][r][col] = (r < N_STAGES && r <= k) ? tresp[(a * N_STAGES + (k - r)) * 3 + c] : 0.0;
    }
    __syncwarp();
    double acc_out = 0.0;
    for (int inst = blockIdx.x; inst < n_inst; inst += gridDim.x) {
        for (int rep = 0; rep < reps; ++rep) {
            const double *Qi = Q + (size_t)inst * N_STAGES * 81;
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) {
                    // W = blockdiag(Q_k^(ab)) A^(b)'   (3N x 24)
                    for (int e = lane; e < KDIM * HP; e += 32) {
                        const int row = e / HP, j = e % HP, k = row / 3, c = row % 3;
                        const double *q = Qi + k * 81 + (a * 3 + b) * 9 + c * 3;
                        S.W[row][j] = q[0] * S.A[b][j][3 * k] + q[1] * S.A[b][j][3 * k + 1] + q[2] * S.A[b][j][3 * k + 2];
                    }
                    __syncwarp();
                    // H^(ab) = A^(a) W on DMMA: 3 x 3 tiles, 15 k-steps
                    for (int ti = 0; ti < 3; ++ti)
                        for (int tj = 0; tj < 3; ++tj) {
                            double c0 = 0.0, c1 = 0.0;
#pragma unroll
                            for (int ks = 0; ks < KDIM / 4; ++ks) {
                                const double av = S.A[a][8 * ti + (lane >> 2)][4 * ks + (lane & 3)];
                                const double bv = S.W[4 * ks + (lane & 3)][8 * tj + (lane >> 2)];
                                dmma(c0, c1, av, bv);
                            }
                            const int r = 8 * ti + (lane >> 2), cc = 8 * tj + 2 * (lane & 3);
                            S.H[24 * a + r][24 * b + cc] = c0 + (a == b && r == cc ? 1.0 : 0.0);
                            S.H[24 * a + r][24 * b + cc + 1] = c1 + (a == b && r == cc + 1 ? 1.0 : 0.0);
                        }
                    __syncwarp();
                }
            if (do_chol) {
                // right-looking blocked Cholesky of the 72 x 72 (padded) matrix, 8 x 8 blocks
                for (int kb = 0; kb < 9; ++kb) {
                    const int o = 8 * kb;
                    // diagonal block: scalar Cholesky (lane j owns column j of the block)
                    for (int j = 0; j < 8; ++j) {
                        double d = S.H[o + j][o + j];
                        d = d > 1e-300 ? sqrt(d) : 1.0;
                        __syncwarp();
                        if (lane == 0) S.H[o + j][o + j] = d;
                        __syncwarp();
                        if (lane > j && lane < 8) S.H[o + lane][o + j] /= d;
                        __syncwarp();
                        if (lane > j && lane < 8)
                            for (int c = j + 1; c <= lane; ++c) S.H[o + lane][o + c] -= S.H[o + lane][o + j] * S.H[o + c][o + j];
                        __syncwarp();
                    }
                    // panel below: solve X L' = B, one row per lane
                    for (int r = o + 8 + lane; r < 72; r += 32)
                        for (int j = 0; j < 8; ++j) {
                            double v = S.H[r][o + j];
                            for (int c = 0; c < j; ++c) v -= S.H[r][o + c] * S.H[o + j][o + c];
                            S.H[r][o + j] = v / S.H[o + j][o + j];
                        }
                    __syncwarp();
                    // trailing update H22 -= L21 L21' on DMMA (lower tiles only)
                    for (int ti = kb + 1; ti < 9; ++ti)
                        for (int tj = kb + 1; tj <= ti; ++tj) {
                            const int r = 8 * ti + (lane >> 2), cc = 8 * tj + 2 * (lane & 3);
                            double c0 = S.H[r][cc], c1 = S.H[r][cc + 1];
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) {
                                const double av = -S.H[8 * ti + (lane >> 2)][o + 4 * ks + (lane & 3)];
                                const double bv = S.H[8 * tj + (lane >> 2)][o + 4 * ks + (lane & 3)];
                                dmma(c0, c1, av, bv);
                            }
                            S.H[r][cc] = c0, S.H[r][cc + 1] = c1;
                        }
                    __syncwarp();
                }
            }
            acc_out += S.H[lane][lane] + S.H[40 + lane][lane];
        }
    }
    out[blockIdx.x * 32 + lane] = acc_out;
}

int main(int argc, char **argv) {
    const int n_inst = argc > 1 ? atoi(argv[1]) : 8192, reps = argc > 2 ? atoi(argv[2]) : 4;
    int dev = 0, n_sm = 0, khz = 0;
    cudaSetDevice(dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    // impulse responses of the reference dynamics (tau from config/mpc_parameters.yaml, dt = 0.05):
    // any decaying 3-vectors exercise the same arithmetic
    std::vector<double> t(3 * N_STAGES * 3);
    for (int a = 0; a < 3; ++a)
        for (int l = 0; l < N_STAGES; ++l) {
            t[(a * N_STAGES + l) * 3 + 0] = 1e-3 * (l + 1) * (l + 1);
            t[(a * N_STAGES + l) * 3 + 1] = 2e-2 * (l + 1);
            t[(a * N_STAGES + l) * 3 + 2] = 0.3 / (1.0 + 0.3 * l);
        }
    std::vector<double> Q((size_t)n_inst * N_STAGES * 81);
    srand(1);
    for (size_t i = 0; i < Q.size(); ++i) Q[i] = (rand() % 1000) * 1e-3;
    for (int i = 0; i < n_inst; ++i) // symmetric positive blocks on the diagonal pairs
        for (int k = 0; k < N_STAGES; ++k)
            for (int a = 0; a < 3; ++a)
                for (int c = 0; c < 3; ++c) Q[((size_t)i * N_STAGES + k) * 81 + (a * 3 + a) * 9 + c * 3 + c] += 50.0;
    double *dt, *dQ, *dout;
    cudaMalloc(&dt, t.size() * 8);
    cudaMalloc(&dQ, Q.size() * 8);
    cudaMemcpy(dt, t.data(), t.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dQ, Q.data(), Q.size() * 8, cudaMemcpyHostToDevice);
    const int smem = (int)sizeof(WarpSmem);
    cudaFuncSetAttribute(condensed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, condensed_kernel, 32, smem);
    const int grid = n_sm * per_sm;
    cudaMalloc(&dout, (size_t)grid * 32 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("{\"what\": \"condensed Newton system on FP64 DMMA, one warp per instance\", \"smem_per_warp_bytes\": %d, "
           "\"warps_per_sm\": %d, \"n_sm\": %d, \"instances\": %d, \"reps\": %d",
           smem, per_sm, n_sm, n_inst, reps);
    for (int chol = 0; chol < 2; ++chol) {
        condensed_kernel<<<grid, 32, smem>>>(dt, dQ, dout, n_inst, 1, chol);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        condensed_kernel<<<grid, 32, smem>>>(dt, dQ, dout, n_inst, reps, chol);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double ns_per = ms * 1e6 / ((double)n_inst * reps);
        const double cyc_per_sm = ns_per * n_sm * (khz * 1e-6);
        // useful flop of the nine 20x60x20 products (dense count, zeros of the triangular A included)
        const double gemm_tflops = 9.0 * 2 * 20 * 60 * 20 * n_inst * reps / (ms * 1e-3) / 1e12;
        printf(", \"%s\": {\"ms\": %.3f, \"ns_per_instance\": %.2f, \"cycles_per_instance_per_sm\": %.0f, "
               "\"gemm_useful_tflops\": %.2f}",
               chol ? "gemm_plus_cholesky" : "gemm_only", ms, ns_per, cyc_per_sm, gemm_tflops);
    }
    cudaError_t err = cudaGetLastError();
    printf(", \"cuda_error\": \"%s\"}\n", cudaGetErrorString(err));
    return 0;
}
