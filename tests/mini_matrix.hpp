// Test-only stand-in for the two Eigen types the adapter touches (Eigen3 is not installed in the build image).
#pragma once
#include <vector>
namespace dftpav_adapter {
class MatrixXd {
 public:
  MatrixXd() : r_(0), c_(0) {}
  MatrixXd(int r, int c) : r_(r), c_(c), v_((size_t)r * c, 0.0) {}
  int rows() const { return r_; }
  int cols() const { return c_; }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
  double& operator()(int i, int j) { return v_[(size_t)j * r_ + i]; }   // column-major like Eigen
  double operator()(int i, int j) const { return v_[(size_t)j * r_ + i]; }
 private:
  int r_, c_;
  std::vector<double> v_;
};
class VectorXd {
 public:
  VectorXd() {}
  explicit VectorXd(int n) : v_(n, 0.0) {}
  int size() const { return (int)v_.size(); }
  double& operator()(int i) { return v_[i]; }
  double operator()(int i) const { return v_[i]; }
 private:
  std::vector<double> v_;
};
}  // namespace dftpav_adapter
