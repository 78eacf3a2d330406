// cswap_host.cpp -- C++ caller of the host mirror (include/qipb200.hpp): a controlled swap
// written with MatrixOp constructors, run on the B200, printed like the reference's README
// example.  Build: see __graft_entry__.build() (g++ -Iinclude ... -Lrustqip_b200 -lqipb200).
#include <cmath>
#include <cstdio>

#include "qipb200.hpp"

int main() {
  using namespace qip;
  typedef std::complex<double> C;
  try {
    Context ctx(0);
    const size_t n = 7;
    const double s = std::sqrt(0.5);
    std::vector<MatrixOp<double>> ops;
    ops.push_back(make_matrix_op<double>({0}, {C(s), C(s), C(s), C(-s)}));                 // H(q)
    for (uint64_t i = 0; i < 3; ++i)                                                        // Fredkin per qubit pair
      ops.push_back(make_control_op<double>({0}, make_swap_op<double>({1 + i}, {4 + i})));
    ops.push_back(make_matrix_op<double>({0}, {C(s), C(s), C(s), C(-s)}));                 // H(q)
    B200State<double> st(ctx, n);
    st.set_basis(4);  // rb = 0b001
    st.apply_all(ops);
    std::vector<double> p = st.measure_probs({0});
    std::vector<C> psi = st.into_state();
    std::printf("P(q=0)=%.3f P(q=1)=%.3f |psi|^2=%.12f\n", p[0], p[1], st.prob_magnitude());
    for (size_t i = 0; i < psi.size(); ++i)
      if (std::abs(psi[i]) > 1e-9) std::printf("  amp[%zu] = %+.3f%+.3fi\n", i, psi[i].real(), psi[i].imag());
    return (std::fabs(p[0] - 0.5) < 1e-12 && std::fabs(psi[4].real() - 0.5) < 1e-12) ? 0 : 1;
  } catch (const CircuitError &e) {
    std::fprintf(stderr, "CircuitError(%d): %s\n", e.status, e.what());
    return 2;
  }
}
