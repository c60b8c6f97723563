// rcs::common::Pose / RPY for the B200 backend's compiled interface (rcs_b200._core): the value semantics of
// /root/reference/include/rcs/Pose.h:23-130 and src/rcs/Pose.cpp without Eigen (this image has none): translation +
// unit quaternion stored (x, y, z, w) like Eigen's coeffs(), every constructor normalises except the one from a bare
// rotation matrix, eulerAngles(2, 1, 0) range convention for rotation_rpy, L1 norm in is_close.
#pragma once
#include <array>
#include <cmath>
#include <sstream>
#include <string>

namespace rcs {
namespace common {

using Vec3 = std::array<double, 3>;
using Vec4 = std::array<double, 4>;   // x y z w
using Mat3 = std::array<double, 9>;   // row-major
using Mat4 = std::array<double, 16>;  // row-major

inline Vec3 IdentityTranslation() { return {0, 0, 0}; }
inline Mat3 IdentityRotMatrix() { return {1, 0, 0, 0, 1, 0, 0, 0, 1}; }
inline Vec4 IdentityRotQuatVec() { return {0, 0, 0, 1}; }
inline Mat4 FrankaHandTCPOffset() {  // Pose.cpp:11-15
  return {0.707, 0.707, 0, 0, -0.707, 0.707, 0, 0, 0, 0, 1, 0.1034, 0, 0, 0, 1};
}

namespace detail {
inline Vec4 qnormalized(Vec4 q) {
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n > 0) for (auto& v : q) v /= n;
  return q;
}
inline Vec4 qmul(const Vec4& a, const Vec4& b) {
  return {a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1], a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2],
          a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0], a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2]};
}
inline Vec3 qrot(const Vec4& q, const Vec3& v) {
  const double ux = 2 * (q[1] * v[2] - q[2] * v[1]), uy = 2 * (q[2] * v[0] - q[0] * v[2]), uz = 2 * (q[0] * v[1] - q[1] * v[0]);
  return {v[0] + q[3] * ux + (q[1] * uz - q[2] * uy), v[1] + q[3] * uy + (q[2] * ux - q[0] * uz), v[2] + q[3] * uz + (q[0] * uy - q[1] * ux)};
}
inline Mat3 qmat(const Vec4& q) {
  const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3], txx = tx * q[0], txy = ty * q[0], txz = tz * q[0], tyy = ty * q[1],
               tyz = tz * q[1], tzz = tz * q[2];
  return {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)};
}
inline Vec4 qfrommat(const Mat3& m) {  // Eigen::Quaterniond(Matrix3d)
  Vec4 q;
  double t = m[0] + m[4] + m[8];
  if (t > 0) {
    t = std::sqrt(t + 1);
    q[3] = 0.5 * t; t = 0.5 / t;
    q[0] = (m[7] - m[5]) * t; q[1] = (m[2] - m[6]) * t; q[2] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1);
    q[i] = 0.5 * t; t = 0.5 / t;
    q[3] = (m[3 * k + j] - m[3 * j + k]) * t;
    q[j] = (m[3 * j + i] + m[3 * i + j]) * t;
    q[k] = (m[3 * k + i] + m[3 * i + k]) * t;
  }
  return q;
}
inline Vec4 qfromrpy(double roll, double pitch, double yaw) {  // Rz(yaw) Ry(pitch) Rx(roll), Pose.h:37-43
  const double cr = std::cos(roll / 2), sr = std::sin(roll / 2), cp = std::cos(pitch / 2), sp = std::sin(pitch / 2), cy = std::cos(yaw / 2),
               sy = std::sin(yaw / 2);
  return {sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy};
}
// polar factor of a 3x3 matrix by Newton iteration X <- (X + X^-T) / 2 (what Eigen's Affine3d::rotation() returns via SVD)
inline Mat3 polar(Mat3 x) {
  for (int it = 0; it < 50; it++) {
    const double c00 = x[4] * x[8] - x[5] * x[7], c01 = x[5] * x[6] - x[3] * x[8], c02 = x[3] * x[7] - x[4] * x[6];
    const double det = x[0] * c00 + x[1] * c01 + x[2] * c02;
    if (std::fabs(det) < 1e-300) break;
    const Mat3 invT = {c00 / det, c01 / det, c02 / det,
                       (x[2] * x[7] - x[1] * x[8]) / det, (x[0] * x[8] - x[2] * x[6]) / det, (x[1] * x[6] - x[0] * x[7]) / det,
                       (x[1] * x[5] - x[2] * x[4]) / det, (x[2] * x[3] - x[0] * x[5]) / det, (x[0] * x[4] - x[1] * x[3]) / det};
    double diff = 0;
    for (int i = 0; i < 9; i++) { const double v = 0.5 * (x[i] + invT[i]); diff += std::fabs(v - x[i]); x[i] = v; }
    if (diff < 1e-15) break;
  }
  return x;
}
}  // namespace detail

struct RPY {
  double roll = 0, pitch = 0, yaw = 0;
  RPY() = default;
  RPY(double r, double p, double y) : roll(r), pitch(p), yaw(y) {}
  explicit RPY(const Vec3& v) : roll(v[0]), pitch(v[1]), yaw(v[2]) {}
  Vec4 as_quaternion_vector() const { return detail::qfromrpy(roll, pitch, yaw); }
  Mat3 rotation_matrix() const { return detail::qmat(as_quaternion_vector()); }
  Vec3 as_vector() const { return {roll, pitch, yaw}; }
  bool is_close(const RPY& o, double eps = 1e-8) const {
    return std::fabs(roll - o.roll) < eps && std::fabs(pitch - o.pitch) < eps && std::fabs(yaw - o.yaw) < eps;
  }
  RPY operator+(const RPY& o) const { return RPY(roll + o.roll, pitch + o.pitch, yaw + o.yaw); }
  std::string str() const { std::ostringstream s; s << "RPY(" << roll << ", " << pitch << ", " << yaw << ")"; return s.str(); }
};

class Pose {
 public:
  Pose() : t_{0, 0, 0}, q_{0, 0, 0, 1} {}
  explicit Pose(const Mat4& m) : t_{m[3], m[7], m[11]} {
    q_ = detail::qnormalized(detail::qfrommat(detail::polar({m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10]})));
  }
  Pose(const Mat3& rotation, const Vec3& translation) : t_(translation), q_(detail::qnormalized(detail::qfrommat(rotation))) {}
  Pose(const Vec4& quaternion, const Vec3& translation) : t_(translation), q_(detail::qnormalized(quaternion)) {}
  Pose(const RPY& rpy, const Vec3& translation) : t_(translation), q_(detail::qnormalized(rpy.as_quaternion_vector())) {}
  static Pose from_rpy_vector(const Vec3& rpy, const Vec3& translation) { return Pose(RPY(rpy), translation); }
  static Pose from_translation(const Vec3& translation) { Pose p; p.t_ = translation; return p; }
  explicit Pose(const Vec4& quaternion) : t_{0, 0, 0}, q_(detail::qnormalized(quaternion)) {}
  explicit Pose(const RPY& rpy) : t_{0, 0, 0}, q_(detail::qnormalized(rpy.as_quaternion_vector())) {}
  explicit Pose(const Mat3& rotation) : t_{0, 0, 0}, q_(detail::qfrommat(rotation)) {}  // not normalised (Pose.cpp:94-97)

  Vec3 translation() const { return t_; }
  Vec4 rotation_q() const { return q_; }
  Mat3 rotation_m() const { return detail::qmat(q_); }
  Mat4 pose_matrix() const {
    const Mat3 r = rotation_m();
    return {r[0], r[1], r[2], t_[0], r[3], r[4], r[5], t_[1], r[6], r[7], r[8], t_[2], 0, 0, 0, 1};
  }
  RPY rotation_rpy() const {  // Eigen eulerAngles(2, 1, 0)
    const Mat3 m = rotation_m();
    double r0 = std::atan2(m[3], m[0]), r1;
    const double c2 = std::sqrt(m[8] * m[8] + m[7] * m[7]);
    if (r0 < 0) { r0 += M_PI; r1 = std::atan2(-m[6], -c2); } else r1 = std::atan2(-m[6], c2);
    const double s1 = std::sin(r0), c1 = std::cos(r0);
    const double r2 = std::atan2(s1 * m[2] - c1 * m[5], c1 * m[4] - s1 * m[1]);
    return RPY(r2, r1, r0);
  }
  std::array<double, 6> xyzrpy() const { const RPY r = rotation_rpy(); return {t_[0], t_[1], t_[2], r.roll, r.pitch, r.yaw}; }
  Pose inverse() const {
    const Vec4 c = {-q_[0], -q_[1], -q_[2], q_[3]};
    const Vec3 r = detail::qrot(c, t_);
    Pose p; p.t_ = {-r[0], -r[1], -r[2]}; p.q_ = detail::qnormalized(c); return p;
  }
  Pose operator*(const Pose& o) const {
    const Vec3 r = detail::qrot(q_, o.t_);
    Pose p; p.t_ = {r[0] + t_[0], r[1] + t_[1], r[2] + t_[2]}; p.q_ = detail::qnormalized(detail::qmul(q_, o.q_)); return p;
  }
  double total_angle() const { return angular_distance(q_, {0, 0, 0, 1}); }
  Pose limit_rotation_angle(double max_angle) const {
    const double a = total_angle();
    if (!(a > max_angle && max_angle >= 0)) return *this;
    Pose p; p.t_ = t_; p.q_ = detail::qnormalized(slerp({0, 0, 0, 1}, max_angle / a, q_)); return p;
  }
  Pose limit_translation_length(double max_length) const {
    const double n = std::sqrt(t_[0] * t_[0] + t_[1] * t_[1] + t_[2] * t_[2]);
    if (!(n > max_length && max_length >= 0)) return *this;
    Pose p; p.q_ = q_; p.t_ = {t_[0] * max_length / n, t_[1] * max_length / n, t_[2] * max_length / n}; return p;
  }
  Pose interpolate(const Pose& dest, double progress) const {
    if (progress > 1) progress = 1;
    Pose p;
    for (int i = 0; i < 3; i++) p.t_[i] = t_[i] + (dest.t_[i] - t_[i]) * progress;
    p.q_ = detail::qnormalized(slerp(q_, progress, dest.q_));
    return p;
  }
  bool is_close(const Pose& o, double eps_r = 1e-8, double eps_t = 1e-8) const {  // L1 norm on the translation (Pose.cpp:208-211)
    return std::fabs(t_[0] - o.t_[0]) + std::fabs(t_[1] - o.t_[1]) + std::fabs(t_[2] - o.t_[2]) < eps_t && angular_distance(q_, o.q_) < eps_r;
  }
  std::string str() const {
    std::ostringstream s;
    s << "Pose(translation=[" << t_[0] << ", " << t_[1] << ", " << t_[2] << "], quaternion=[" << q_[0] << ", " << q_[1] << ", " << q_[2] << ", "
      << q_[3] << "])";
    return s.str();
  }
  static Pose from_tq(const Vec3& t, const Vec4& q) { Pose p; p.t_ = t; p.q_ = q; return p; }  // as stored, not normalised

 private:
  static double angular_distance(const Vec4& a, const Vec4& b) {  // Eigen: 2 atan2(|vec(d)|, |w(d)|), d = a * conj(b)
    const Vec4 d = detail::qmul(a, {-b[0], -b[1], -b[2], b[3]});
    return 2 * std::atan2(std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), std::fabs(d[3]));
  }
  static Vec4 slerp(const Vec4& a, double t, const Vec4& b) {  // Eigen QuaternionBase::slerp
    const double d = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3], ad = std::fabs(d);
    double s0, s1;
    if (ad >= 1.0 - 2.220446049250313e-16) { s0 = 1 - t; s1 = t; }
    else { const double th = std::acos(ad), st = std::sin(th); s0 = std::sin((1 - t) * th) / st; s1 = std::sin(t * th) / st; }
    if (d < 0) s1 = -s1;
    return {s0 * a[0] + s1 * b[0], s0 * a[1] + s1 * b[1], s0 * a[2] + s1 * b[2], s0 * a[3] + s1 * b[3]};
  }
  Vec3 t_;
  Vec4 q_;
};

}  // namespace common
}  // namespace rcs
