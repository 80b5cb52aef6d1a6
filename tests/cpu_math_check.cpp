// Host build of sqrtlm-slam_b200/csrc/sqrtba_math.cuh (the __host__ __device__ arithmetic the CUDA kernels use),
// so the per-observation formulas can be compared against the oracle on a machine without a GPU.
#include "../sqrtlm-slam_b200/csrc/sqrtba_math.cuh"

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
}
