// pcaone_b200 host — shared types of the C++ front-end (the reference's Common.hpp /
// Logger.hpp / Timer.hpp roles, /root/reference/src/Common.hpp:16-68, Logger.hpp:20-100).
// The reference's matrices are Eigen::MatrixXd; the GPU path only needs a column-major
// buffer with the same memory layout as Eigen::MatrixXd::data(), so the drop-in keeps the
// reference's type NAMES (Mat2D, Mat1D) on a minimal owning container.
#pragma once
#include <chrono>
#include <cstdint>
#include <ctime>
#include <fstream>
#include <iostream>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace pcaone_host {

using uint = unsigned int;
using uint64 = uint64_t;

struct Mat2D {  // column-major, like Eigen::MatrixXd
  uint64 nrow = 0, ncol = 0;
  std::vector<double> v;
  Mat2D() = default;
  Mat2D(uint64 r, uint64 c) : nrow(r), ncol(c), v(r * c, 0.0) {}
  void resize(uint64 r, uint64 c) {
    nrow = r;
    ncol = c;
    v.assign(r * c, 0.0);
  }
  uint64 rows() const { return nrow; }
  uint64 cols() const { return ncol; }
  double* data() { return v.data(); }
  const double* data() const { return v.data(); }
  double& operator()(uint64 i, uint64 j) { return v[j * nrow + i]; }
  double operator()(uint64 i, uint64 j) const { return v[j * nrow + i]; }
};

struct Mat1D {
  std::vector<double> v;
  Mat1D() = default;
  explicit Mat1D(uint64 n) : v(n, 0.0) {}
  void resize(uint64 n) { v.assign(n, 0.0); }
  uint64 size() const { return v.size(); }
  double* data() { return v.data(); }
  const double* data() const { return v.data(); }
  double& operator()(uint64 i) { return v[i]; }
  double operator()(uint64 i) const { return v[i]; }
};

// wall-clock helper with the three calls the reference's Timer offers (Timer.hpp)
class Timer {
  using clk = std::chrono::steady_clock;
  clk::time_point start_ = clk::now(), mark_ = clk::now();

 public:
  void clock() { mark_ = clk::now(); }
  double reltime() const { return std::chrono::duration<double>(clk::now() - mark_).count(); }
  double abstime() const { return std::chrono::duration<double>(clk::now() - start_).count(); }
  std::string date() const {
    std::time_t t = std::time(nullptr);
    char buf[64];
    std::strftime(buf, sizeof buf, "[%Y-%m-%d %H:%M:%S]", std::localtime(&t));
    return buf;
  }
};

// log to <out>.log and optionally the screen; error() logs and throws std::runtime_error
// exactly like cao.error (Logger.hpp:85-94)
class Logger {
 public:
  std::ofstream file;
  bool is_screen = false;
  // ranks > 0 of an in-process multi-GPU job keep quiet (rank 0 speaks for the job)
  static inline thread_local bool muted = false;

  template <class... A>
  void print(const A&... a) {
    std::ostringstream os;
    join(os, a...);
    emit(os.str(), false);
  }
  template <class... A>
  void warn(const A&... a) {
    std::ostringstream os;
    os << "WARNING: ";
    join(os, a...);
    emit(os.str(), true);
  }
  template <class... A>
  [[noreturn]] void error(const A&... a) {
    std::ostringstream os;
    os << "ERROR: ";
    join(os, a...);
    emit(os.str(), true);
    throw std::runtime_error(os.str());
  }

 private:
  std::mutex mu_;
  template <class T, class... A>
  static void join(std::ostringstream& os, const T& t, const A&... a) {
    os << t;
    if constexpr (sizeof...(a) > 0) {
      os << ' ';
      join(os, a...);
    }
  }
  void emit(const std::string& s, bool err) {
    if (muted && !err) return;
    std::lock_guard<std::mutex> lk(mu_);
    if (file.is_open()) file << s << std::endl;
    if (err)
      std::cerr << s << std::endl;
    else if (is_screen)
      std::cout << s << std::endl;
  }
};

extern Logger cao;
extern Timer tick;

}  // namespace pcaone_host
