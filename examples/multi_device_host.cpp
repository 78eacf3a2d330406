// multi_device_host.cpp -- ONE process driving every visible B200 through one context
// (qipb200_init_multi): the shape a Rust `B200Builder` replacing LocalBuilder::calculate_state_with_init
// (qip/src/builder.rs:400-519) has -- no ranks, no IPC handles in the caller.  A GHZ-style circuit whose
// Hadamard sits on qubit 0 (held by the device index) and whose CNOT chain crosses every shard boundary:
// the expected state is (|0..0> + |1..1>)/sqrt(2), measured and collapsed across the devices.
// Build: see __graft_entry__.build().  Usage: multi_device_host [n_devices] (default: largest power of two visible)
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include <cuda_runtime_api.h>

#include "qipb200.hpp"

int main(int argc, char **argv) {
  using namespace qip;
  typedef std::complex<double> C;
  int visible = 0;
  if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1) {
    std::fprintf(stderr, "no CUDA device: libqipb200 has no CPU path\n");
    return 2;
  }
  int g = 1;
  while (2 * g <= visible) g *= 2;
  if (argc > 1) g = std::atoi(argv[1]);
  try {
    std::vector<int> devs;
    for (int i = 0; i < g; ++i) devs.push_back(i);
    Context ctx(devs);
    const size_t n = 20;
    const double s = std::sqrt(0.5);
    std::vector<MatrixOp<double>> ops;
    ops.push_back(make_matrix_op<double>({0}, {C(s), C(s), C(s), C(-s)}));  // H on qubit 0 (the top index bit)
    for (uint64_t q = 0; q + 1 < n; ++q)                                     // CNOT chain q -> q+1
      ops.push_back(make_control_op<double>({q}, make_matrix_op<double>({q + 1}, {C(0), C(1), C(1), C(0)})));
    B200State<double> st(ctx, n);
    st.set_basis(0);
    st.apply_all(ops);
    std::vector<double> p = st.measure_probs({0, n - 1});
    const double nrm = st.prob_magnitude();
    std::vector<C> psi = st.into_state();
    const size_t last = (size_t(1) << n) - 1;
    std::printf("%d device(s): |psi|^2=%.12f  P(00)=%.3f P(01)=%.3f P(10)=%.3f P(11)=%.3f  amp[0]=%+.6f amp[2^n-1]=%+.6f\n", g, nrm,
                p[0], p[1], p[2], p[3], psi[0].real(), psi[last].real());
    const uint64_t m = st.soft_measure({0}, 0.75);  // the draw 0.75 falls into the |1..1> half
    st.collapse({0}, m, 0.5);
    std::vector<C> post = st.into_state();
    std::printf("measured qubit 0 = %llu, post-measurement amp[2^n-1] = %+.6f\n", (unsigned long long)m, post[last].real());
    const bool ok = std::fabs(nrm - 1.0) < 1e-12 && std::fabs(p[0] - 0.5) < 1e-12 && std::fabs(p[3] - 0.5) < 1e-12 &&
                    std::fabs(psi[0].real() - s) < 1e-12 && std::fabs(psi[last].real() - s) < 1e-12 && m == 1 &&
                    std::fabs(post[last].real() - 1.0) < 1e-12 && std::abs(post[0]) < 1e-12;
    return ok ? 0 : 1;
  } catch (const CircuitError &e) {
    std::fprintf(stderr, "CircuitError(%d): %s\n", e.status, e.what());
    return 2;
  }
}
