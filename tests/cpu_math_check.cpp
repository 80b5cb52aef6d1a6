// Host build of sqrtlm-slam_b200/csrc/sqrtba_math.cuh (the __host__ __device__ arithmetic the CUDA kernels use),
// so the per-observation formulas can be compared against the oracle on a machine without a GPU.
#include "../sqrtlm-slam_b200/csrc/sqrtba_math.cuh"
#include "../sqrtlm-slam_b200/csrc/sqrtba_sim3.cuh"

extern "C" {
void mc_obs(const double* pose7, const double* X, const double* cam, const float* meas4, double* e, double* Jp,
            double* Jl, int* depth_pos) {
  double R[9], Xc[3];
  sqrtba::quat_to_R(pose7 + 3, R);
  sqrtba::transform(R, pose7, X, Xc);
  const bool stereo = !(meas4[2] < 0.0f);
  sqrtba::reproj_error(Xc, meas4[0], meas4[1], meas4[2], cam, stereo, e);
  sqrtba::reproj_jacobians(R, Xc, cam, stereo, Jp, Jl);
  *depth_pos = Xc[2] > 0.0;
}
void mc_huber(double c, double delta, double* rho0, double* rho1) { sqrtba::huber(c, delta, sqrtba::huber_dsqr(delta), rho0, rho1); }
void mc_oplus(double* pose7, const double* xi) { sqrtba::pose_oplus(pose7, xi); }
int mc_spd6_inverse(const double* A, double* Ai) { return sqrtba::spd6_inverse(A, Ai) ? 1 : 0; }
void mc_sim3_exp(const double* u7, double* out8) { sqrtba::sim3_exp(u7, out8); }
void mc_sim3_log(const double* s8, double* out7) { sqrtba::sim3_log(s8, out7); }
void mc_sim3_oplus(double* est8, const double* u7, int fix_scale) { sqrtba::sim3_oplus(est8, u7, fix_scale != 0); }
void mc_sim3_edge_linearize(const double* c8, const double* a8, const double* b8, int f1, int f2, int fix_scale, double* Ji,
                            double* Jj) {
  sqrtba::sim3_edge_linearize(c8, a8, b8, f1 != 0, f2 != 0, fix_scale != 0, Ji, Jj);
}
void mc_sim3_edge_error(const double* c8, const double* a8, const double* b8, double* err7) { sqrtba::sim3_edge_error(c8, a8, b8, err7); }
}
