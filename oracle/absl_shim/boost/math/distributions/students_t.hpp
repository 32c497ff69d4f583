// Build scaffolding for oracle/_ref ONLY: boost::math::students_t with quantile(complement(dist, p)), which is all
// cc/mcts/tree.cc:17-38 uses (two-sided critical values for its LCB table).  CDF through the regularised incomplete beta
// function (Lentz continued fraction), quantile by bisection to 1e-12: the published definitions, not boost's code.
#pragma once
#include <cmath>
namespace boost {
namespace math {
namespace shim {
inline double BetaCf(double a, double b, double x) {
  const double tiny = 1e-300;
  double qab = a + b, qap = a + 1, qam = a - 1, c = 1, d = 1 - qab * x / qap;
  if (std::fabs(d) < tiny) d = tiny;
  d = 1 / d;
  double h = d;
  for (int m = 1; m <= 500; ++m) {
    const int m2 = 2 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1 + aa * d; if (std::fabs(d) < tiny) d = tiny;
    c = 1 + aa / c; if (std::fabs(c) < tiny) c = tiny;
    d = 1 / d; h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1 + aa * d; if (std::fabs(d) < tiny) d = tiny;
    c = 1 + aa / c; if (std::fabs(c) < tiny) c = tiny;
    d = 1 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1) < 1e-15) break;
  }
  return h;
}
inline double IncBeta(double a, double b, double x) {  // I_x(a, b)
  if (x <= 0) return 0;
  if (x >= 1) return 1;
  const double bt = std::exp(std::lgamma(a + b) - std::lgamma(a) - std::lgamma(b) + a * std::log(x) + b * std::log(1 - x));
  return x < (a + 1) / (a + b + 2) ? bt * BetaCf(a, b, x) / a : 1 - bt * BetaCf(b, a, 1 - x) / b;
}
inline double TUpperTail(double t, double v) {  // P(T > t), t >= 0
  return 0.5 * IncBeta(v / 2, 0.5, v / (v + t * t));
}
}  // namespace shim
class students_t {
 public:
  explicit students_t(double v) : v_(v) {}
  double degrees_of_freedom() const { return v_; }

 private:
  double v_;
};
template <typename D>
struct complemented {
  D dist;
  double p;
};
template <typename D>
complemented<D> complement(const D& d, double p) { return {d, p}; }
inline double quantile(const complemented<students_t>& c) {  // t with P(T > t) = p
  const double p = c.p, v = c.dist.degrees_of_freedom();
  if (p >= 0.5) return p == 0.5 ? 0.0 : -quantile(complemented<students_t>{c.dist, 1 - p});
  double lo = 0, hi = 1;
  while (shim::TUpperTail(hi, v) > p && hi < 1e12) hi *= 2;
  for (int i = 0; i < 200; ++i) {
    const double mid = 0.5 * (lo + hi);
    if (shim::TUpperTail(mid, v) > p) lo = mid; else hi = mid;
  }
  return 0.5 * (lo + hi);
}
}  // namespace math
}  // namespace boost
