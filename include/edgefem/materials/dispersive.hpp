// Frequency-dependent material models (reference: include/edgefem/materials/dispersive.hpp).
// e^{+j w t} convention: loss gives Im(eps) < 0.  Host evaluation below is used for port
// propagation constants; inside assemble_maxwell the models are evaluated on the GPU from the
// parameters exposed by describe().
#pragma once
#include <array>
#include <complex>
#include <memory>
#include <stdexcept>
#include <vector>

namespace edgefem {
namespace materials {

/// Parameter view used to ship a model to the device (kind matches efb_model_kind).
struct ModelDescription {
  int kind = 0; // 0 = opaque (evaluate on host), 1 Debye, 2 Lorentz, 3 Drude, 4 Drude-Lorentz
  double p0 = 0.0, p1 = 0.0, p2 = 0.0;
  std::vector<std::array<double, 3>> poles; // {delta_eps, omega0, gamma}
};

class DispersiveMaterial {
public:
  virtual ~DispersiveMaterial() = default;
  virtual std::complex<double> eval_eps(double omega) const = 0;
  virtual std::complex<double> eval_mu(double omega) const {
    (void)omega;
    return {1.0, 0.0};
  }
  /// kind 0 => user subclass: the host evaluates eval_eps()/eval_mu() per frequency instead
  virtual ModelDescription describe() const { return {}; }
};

class DebyeMaterial : public DispersiveMaterial {
public:
  DebyeMaterial(double eps_static, double eps_inf, double tau) : es_(eps_static), ei_(eps_inf), tau_(tau) {
    if (tau <= 0.0) throw std::invalid_argument("DebyeMaterial: tau must be positive");
  }
  std::complex<double> eval_eps(double omega) const override {
    return ei_ + (es_ - ei_) / std::complex<double>(1.0, omega * tau_);
  }
  ModelDescription describe() const override { return {1, es_, ei_, tau_, {}}; }
  double eps_static() const { return es_; }
  double eps_inf() const { return ei_; }
  double tau() const { return tau_; }

private:
  double es_, ei_, tau_;
};

class LorentzMaterial : public DispersiveMaterial {
public:
  struct Pole {
    double delta_eps, omega0, gamma;
  };
  LorentzMaterial() : ei_(1.0) {}
  explicit LorentzMaterial(double eps_inf) : ei_(eps_inf) {}
  void add_pole(double delta_eps, double omega0, double gamma) {
    if (omega0 <= 0.0) throw std::invalid_argument("LorentzMaterial: omega0 must be positive");
    if (gamma < 0.0) throw std::invalid_argument("LorentzMaterial: gamma must be non-negative");
    poles_.push_back({delta_eps, omega0, gamma});
  }
  std::complex<double> eval_eps(double omega) const override {
    std::complex<double> eps = ei_;
    const double w2 = omega * omega;
    for (const auto &p : poles_) {
      const double w02 = p.omega0 * p.omega0;
      eps += p.delta_eps * w02 / std::complex<double>(w02 - w2, p.gamma * omega);
    }
    return eps;
  }
  ModelDescription describe() const override {
    ModelDescription d{2, ei_, 0.0, 0.0, {}};
    for (const auto &p : poles_) d.poles.push_back({p.delta_eps, p.omega0, p.gamma});
    return d;
  }
  double eps_inf() const { return ei_; }
  void set_eps_inf(double e) { ei_ = e; }
  const std::vector<Pole> &poles() const { return poles_; }
  size_t num_poles() const { return poles_.size(); }

private:
  double ei_;
  std::vector<Pole> poles_;
};

class DrudeMaterial : public DispersiveMaterial {
public:
  DrudeMaterial(double omega_p, double gamma) : wp_(omega_p), g_(gamma) {
    if (omega_p <= 0.0) throw std::invalid_argument("DrudeMaterial: omega_p must be positive");
    if (gamma < 0.0) throw std::invalid_argument("DrudeMaterial: gamma must be non-negative");
  }
  std::complex<double> eval_eps(double omega) const override {
    if (omega == 0.0) return {-1e30, 0.0};
    return std::complex<double>(1.0, 0.0) - (wp_ * wp_) / std::complex<double>(omega * omega, g_ * omega);
  }
  ModelDescription describe() const override { return {3, wp_, g_, 0.0, {}}; }
  double omega_p() const { return wp_; }
  double gamma() const { return g_; }

private:
  double wp_, g_;
};

class DrudeLorentzMaterial : public DispersiveMaterial {
public:
  struct LorentzPole {
    double delta_eps, omega0, gamma;
  };
  DrudeLorentzMaterial(double eps_inf, double omega_p, double gamma_d) : ei_(eps_inf), wp_(omega_p), gd_(gamma_d) {
    if (omega_p <= 0.0) throw std::invalid_argument("DrudeLorentzMaterial: omega_p must be positive");
    if (gamma_d < 0.0) throw std::invalid_argument("DrudeLorentzMaterial: gamma_d must be non-negative");
  }
  void add_lorentz_pole(double delta_eps, double omega0, double gamma) {
    if (omega0 <= 0.0) throw std::invalid_argument("DrudeLorentzMaterial: omega0 must be positive");
    if (gamma < 0.0) throw std::invalid_argument("DrudeLorentzMaterial: gamma must be non-negative");
    poles_.push_back({delta_eps, omega0, gamma});
  }
  std::complex<double> eval_eps(double omega) const override {
    std::complex<double> eps = ei_;
    if (omega != 0.0)
      eps -= (wp_ * wp_) / std::complex<double>(omega * omega, gd_ * omega);
    else
      eps = {-1e30, 0.0};
    const double w2 = omega * omega;
    for (const auto &p : poles_) {
      const double w02 = p.omega0 * p.omega0;
      eps += p.delta_eps * w02 / std::complex<double>(w02 - w2, p.gamma * omega);
    }
    return eps;
  }
  ModelDescription describe() const override {
    ModelDescription d{4, ei_, wp_, gd_, {}};
    for (const auto &p : poles_) d.poles.push_back({p.delta_eps, p.omega0, p.gamma});
    return d;
  }
  double eps_inf() const { return ei_; }
  double omega_p() const { return wp_; }
  double gamma_d() const { return gd_; }
  const std::vector<LorentzPole> &lorentz_poles() const { return poles_; }

private:
  double ei_, wp_, gd_;
  std::vector<LorentzPole> poles_;
};

} // namespace materials
} // namespace edgefem
