// The C++ interfaces hardware extensions and the Python layer program against, restated for the B200 backend's
// compiled module (rcs_b200._core): /root/reference/include/rcs/Robot.h:24-197 (RobotType, RobotMetaConfig,
// RobotConfig / RobotState, Robot, Gripper) and include/rcs/Kinematics.h:19-26 (Kinematics). Same names, virtual
// signatures and defaults; std::vector<double> / std::array stand in for the Eigen types (no Eigen in this image).
#pragma once
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "Pose.h"

namespace rcs {
namespace common {

using VectorXd = std::vector<double>;

enum RobotType { FR3 = 0, UR5e, SO101, XArm7 };          // Robot.h:19
enum RobotPlatform { SIMULATION = 0, HARDWARE };          // Robot.h:97

struct RobotMetaConfig {  // Robot.h:21-27
  VectorXd q_home;
  int dof;
  std::array<VectorXd, 2> joint_limits;  // low, high
};
inline const RobotMetaConfig& robots_meta_config(RobotType t) {  // Robot.h:28-95
  static const double d2r = M_PI / 180.0;
  static const RobotMetaConfig fr3 = {{0, -M_PI_4, 0, -3 * M_PI_4, 0, M_PI_2, M_PI_4}, 7,
      {VectorXd{-2.3093, -1.5133, -2.4937, -2.7478, -2.4800, 0.8521, -2.6895}, VectorXd{2.3093, 1.5133, 2.4937, -0.4461, 2.4800, 4.2094, 2.6895}}};
  static const RobotMetaConfig ur5e = {{-0.4488354, -2.02711196, 1.64630026, -1.18999615, -1.57079762, -2.01963249}, 6,
      {VectorXd{-2 * M_PI, -2 * M_PI, -M_PI, -2 * M_PI, -2 * M_PI, -2 * M_PI}, VectorXd{2 * M_PI, 2 * M_PI, M_PI, 2 * M_PI, 2 * M_PI, 2 * M_PI}}};
  static const RobotMetaConfig xarm7 = {{0, -45 * d2r, 0, 15 * d2r, 0, -25 * d2r, 0}, 7,
      {VectorXd{-2 * M_PI, -2.094395, -2 * M_PI, -3.92699, -2 * M_PI, -M_PI, -2 * M_PI}, VectorXd{2 * M_PI, 2.059488, 2 * M_PI, 0.191986, 2 * M_PI, 1.692969, 2 * M_PI}}};
  static const RobotMetaConfig so101 = {{-9.40612320177057, -99.66130397967824, 99.9124726477024, 69.96996996996998, -9.095744680851055}, 5,
      {VectorXd{-100, -100, -100, -100, -100}, VectorXd{100, 100, 100, 100, 100}}};
  switch (t) {
    case FR3: return fr3;
    case UR5e: return ur5e;
    case XArm7: return xarm7;
    case SO101: return so101;
  }
  throw std::invalid_argument("unknown robot type");
}

class Kinematics {  // Kinematics.h:19-26
 public:
  virtual ~Kinematics() = default;
  virtual std::optional<VectorXd> inverse(const Pose& pose, const VectorXd& q0, const Pose& tcp_offset = Pose()) = 0;
  virtual Pose forward(const VectorXd& q0, const Pose& tcp_offset) = 0;
};

struct RobotConfig {  // Robot.h:99-107
  RobotType robot_type = FR3;
  RobotPlatform robot_platform = SIMULATION;
  Pose tcp_offset = Pose();
  std::string attachment_site = "attachment_site";
  std::string kinematic_model_path = "assets/scenes/fr3_empty_world/robot.xml";
  virtual ~RobotConfig() = default;
};
struct RobotState { virtual ~RobotState() = default; };
struct GripperConfig { virtual ~GripperConfig() = default; };
struct GripperState { virtual ~GripperState() = default; };

class Robot {  // Robot.h:127-161
 public:
  virtual ~Robot() = default;
  virtual RobotConfig* get_config() = 0;
  virtual RobotState* get_state() = 0;
  virtual Pose get_cartesian_position() = 0;
  virtual void set_joint_position(const VectorXd& q) = 0;
  virtual VectorXd get_joint_position() = 0;
  virtual void move_home() = 0;
  virtual void reset() = 0;
  virtual void close() = 0;
  virtual void set_cartesian_position(const Pose& pose) = 0;
  virtual std::optional<std::shared_ptr<Kinematics>> get_ik() = 0;
  virtual Pose get_base_pose_in_world_coordinates() = 0;
  Pose to_pose_in_robot_coordinates(const Pose& pose_in_world_coordinates) {  // Robot.cpp:5-8
    return get_base_pose_in_world_coordinates().inverse() * pose_in_world_coordinates;
  }
  Pose to_pose_in_world_coordinates(const Pose& pose_in_robot_coordinates) {  // Robot.cpp:10-13
    return get_base_pose_in_world_coordinates() * pose_in_robot_coordinates;
  }
};

class Gripper {  // Robot.h:163-197
 public:
  virtual ~Gripper() = default;
  virtual GripperConfig* get_config() = 0;
  virtual GripperState* get_state() = 0;
  virtual void set_normalized_width(double width, double force = 0) = 0;
  virtual double get_normalized_width() = 0;
  virtual bool is_grasped() = 0;
  virtual void grasp() = 0;
  virtual void open() = 0;
  virtual void shut() = 0;
  virtual void reset() = 0;
  virtual void close() = 0;
};

}  // namespace common
}  // namespace rcs
