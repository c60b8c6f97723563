// rcs_b200._core -- the compiled interface of the B200 backend, shaped like the reference's pybind11 module rcs._core
// (/root/reference/src/pybind/rcs.cpp:186-527): submodules `common` (Pose, RPY, RobotType, RobotMetaConfig, the Kinematics /
// Robot / Gripper interfaces with trampolines so that Python or C++ extensions can subclass them, Pin) and `sim` (SimConfig,
// Sim, SimRobotConfig / State, SimRobot, SimGripperConfig / State, SimGripper). It is host code only: everything it does goes
// through the C ABI of librcsb.so (include/rcsb.h).
//
// Ownership follows the reference: there, Python owns mjModel / mjData and `Sim(mjmdl: int, mjdata: int)` borrows their raw
// addresses (rcs.cpp:493-506, python/rcs/sim/sim.py:47-55). Here Python owns the rcsb_model / rcsb_batch handles
// (rcs_b200.batch.DeviceModel / Batch: device memory lives in torch tensors) and `Sim(model_addr, batch_addr)` borrows them.
// The robot / gripper configuration is compiled into the device model when it is built, so SimRobot / SimGripper bind to the
// batch they are given. With a batch of N environments the single-environment calls of the reference act on every
// environment alike and the getters report environment 0 (the vector API lives in rcs_b200.envs).
#include <pybind11/numpy.h>
#include <pybind11/operators.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdint>
#include <cstring>

#include "../../include/rcs_b200/Robot.h"
#include "../../include/rcsb.h"
#include "rcsb_types.h"

namespace py = pybind11;
using namespace rcs::common;

static void check(int rc) {
  if (rc != 0) throw std::runtime_error(std::string("rcsb error ") + std::to_string(rc) + ": " + rcsb_last_error());
}
template <size_t N>
static std::array<double, N> arr_in(const py::array_t<double, py::array::c_style | py::array::forcecast>& a, const char* what) {
  if ((size_t)a.size() != N) throw py::type_error(std::string(what) + ": wrong number of elements");
  std::array<double, N> o;
  std::memcpy(o.data(), a.data(), N * sizeof(double));
  return o;
}
static py::array_t<double> arr_out(const double* d, std::vector<py::ssize_t> shape) {
  py::array_t<double> a(shape);
  std::memcpy(a.mutable_data(), d, (size_t)a.size() * sizeof(double));
  return a;
}
using NpArr = py::array_t<double, py::array::c_style | py::array::forcecast>;

// ------------------------------------------------------------------ trampolines (rcs.cpp:26-184)
class PyKinematics : public Kinematics {
 public:
  std::optional<VectorXd> inverse(const Pose& pose, const VectorXd& q0, const Pose& tcp_offset) override {
    PYBIND11_OVERRIDE_PURE(std::optional<VectorXd>, Kinematics, inverse, pose, q0, tcp_offset);
  }
  Pose forward(const VectorXd& q0, const Pose& tcp_offset) override { PYBIND11_OVERRIDE_PURE(Pose, Kinematics, forward, q0, tcp_offset); }
};
class PyRobot : public Robot {
 public:
  RobotConfig* get_config() override { PYBIND11_OVERRIDE_PURE(RobotConfig*, Robot, get_config, ); }
  RobotState* get_state() override { PYBIND11_OVERRIDE_PURE(RobotState*, Robot, get_state, ); }
  Pose get_cartesian_position() override { PYBIND11_OVERRIDE_PURE(Pose, Robot, get_cartesian_position, ); }
  void set_joint_position(const VectorXd& q) override { PYBIND11_OVERRIDE_PURE(void, Robot, set_joint_position, q); }
  VectorXd get_joint_position() override { PYBIND11_OVERRIDE_PURE(VectorXd, Robot, get_joint_position, ); }
  void move_home() override { PYBIND11_OVERRIDE_PURE(void, Robot, move_home, ); }
  void reset() override { PYBIND11_OVERRIDE_PURE(void, Robot, reset, ); }
  void close() override { PYBIND11_OVERRIDE_PURE(void, Robot, close, ); }
  void set_cartesian_position(const Pose& pose) override { PYBIND11_OVERRIDE_PURE(void, Robot, set_cartesian_position, pose); }
  std::optional<std::shared_ptr<Kinematics>> get_ik() override {
    PYBIND11_OVERRIDE_PURE(std::optional<std::shared_ptr<Kinematics>>, Robot, get_ik, );
  }
  Pose get_base_pose_in_world_coordinates() override { PYBIND11_OVERRIDE_PURE(Pose, Robot, get_base_pose_in_world_coordinates, ); }
};
class PyGripper : public Gripper {
 public:
  GripperConfig* get_config() override { PYBIND11_OVERRIDE_PURE(GripperConfig*, Gripper, get_config, ); }
  GripperState* get_state() override { PYBIND11_OVERRIDE_PURE(GripperState*, Gripper, get_state, ); }
  void set_normalized_width(double width, double force) override { PYBIND11_OVERRIDE_PURE(void, Gripper, set_normalized_width, width, force); }
  double get_normalized_width() override { PYBIND11_OVERRIDE_PURE(double, Gripper, get_normalized_width, ); }
  bool is_grasped() override { PYBIND11_OVERRIDE_PURE(bool, Gripper, is_grasped, ); }
  void grasp() override { PYBIND11_OVERRIDE_PURE(void, Gripper, grasp, ); }
  void open() override { PYBIND11_OVERRIDE_PURE(void, Gripper, open, ); }
  void shut() override { PYBIND11_OVERRIDE_PURE(void, Gripper, shut, ); }
  void reset() override { PYBIND11_OVERRIDE_PURE(void, Gripper, reset, ); }
  void close() override { PYBIND11_OVERRIDE_PURE(void, Gripper, close, ); }
};

// ------------------------------------------------------------------ sim layer over the C ABI
namespace sim {
struct SimConfig {  // src/sim/sim.h:29-34
  bool async_control = false, realtime = false;
  int frequency = 30, max_convergence_steps = 500;
};
struct BatchInfo { int n = 0, nj = 0, ik_nq = 0, nq = 0, nv = 0, nu = 0, nsr = 0, o_q = 0, o_v = 0, o_ctrl = 0, o_warm = 0, o_tail = 0; };

class Sim {  // src/sim/sim.h:36-78 over a borrowed rcsb_batch
 public:
  Sim(std::uintptr_t model_addr, std::uintptr_t batch_addr) : model((rcsb_model*)model_addr), batch((rcsb_batch*)batch_addr) {
    if (!model || !batch) throw std::invalid_argument("Sim(model_addr, batch_addr): null handle");
    check(rcsb_batch_info(batch, &info.n, &info.nj, &info.ik_nq, &info.nq, &info.nv, &info.nu));
    int nsd, nsi, od, id;
    check(rcsb_model_dims(model, &info.nsr, &nsd, &nsi, &od, &id));
    check(rcsb_model_offsets(model, &info.o_q, &info.o_v, &info.o_ctrl, &info.o_warm, &info.o_tail));
  }
  void step(size_t k) { check(rcsb_sim_step(batch, (int)k)); }
  void step_until_convergence() {
    check(rcsb_sim_step_until_convergence(batch, cfg.max_convergence_steps));
    std::vector<int> si(RCSB_I_TAIL);
    check(rcsb_batch_read_row(batch, 0, nullptr, nullptr, si.data()));
    converged = si[RCSB_I_CONVERGED] != 0;
    if (si[RCSB_I_CONV_STEPS] == cfg.max_convergence_steps) fprintf(stderr, "WARNING: Max convergence steps reached!\n");  // sim.cpp:103-105
  }
  bool is_converged() const { return converged; }
  void reset() { check(rcsb_sim_reset(batch)); }
  bool set_config(const SimConfig& c) { cfg = c; return true; }
  SimConfig get_config() const { return cfg; }
  void row(std::vector<double>& sr, std::vector<int>& si) const {
    sr.resize(info.nsr); si.resize(RCSB_I_TAIL);
    check(rcsb_batch_read_row(batch, 0, sr.data(), nullptr, si.data()));
  }
  // one command for every environment through the host-buffer entry point
  void command(unsigned ops, const VectorXd* joints, const double* gripper) {
    std::vector<double> aj, ag;
    if (joints) {
      if ((int)joints->size() < info.nj) throw std::invalid_argument("joint vector too short");
      aj.resize((size_t)info.n * info.nj);
      for (int e = 0; e < info.n; e++) std::memcpy(&aj[(size_t)e * info.nj], joints->data(), info.nj * sizeof(double));
    }
    if (gripper) ag.assign(info.n, *gripper);
    check(rcsb_batch_run_host(batch, ops, 0, 0, joints ? aj.data() : nullptr, gripper ? ag.data() : nullptr, 0, nullptr, nullptr, nullptr, nullptr));
  }
  rcsb_model* model;
  rcsb_batch* batch;
  BatchInfo info;
  SimConfig cfg;
  bool converged = true;
};

struct SimRobotConfig : RobotConfig {  // src/sim/SimRobot.h:14-47
  double joint_rotational_tolerance = .05 * (M_PI / 180.0), seconds_between_callbacks = 0.1;
  bool trajectory_trace = false;
  std::vector<std::string> arm_collision_geoms{"fr3_link0_collision", "fr3_link1_collision", "fr3_link2_collision", "fr3_link3_collision",
                                               "fr3_link4_collision", "fr3_link5_collision", "fr3_link6_collision", "fr3_link7_collision"};
  std::vector<std::string> joints{"fr3_joint1", "fr3_joint2", "fr3_joint3", "fr3_joint4", "fr3_joint5", "fr3_joint6", "fr3_joint7"};
  std::vector<std::string> actuators = joints;
  std::string base = "base", mjcf_scene_path = "assets/scenes/fr3_empty_world/scene.xml";
  void add_id(const std::string& id) {
    for (auto& s : arm_collision_geoms) s += "_" + id;
    for (auto& s : joints) s += "_" + id;
    for (auto& s : actuators) s += "_" + id;
    attachment_site += "_" + id;
    base += "_" + id;
  }
};
struct SimRobotState : RobotState {  // SimRobot.h:49-57
  VectorXd previous_angles, target_angles;
  Pose inverse_tcp_offset;
  bool ik_success = true, collision = false, is_moving = false, is_arrived = false;
};

class Pin : public Kinematics {  // rcs::common::Pin (src/rcs/Kinematics.cpp:13-81) on the batched CLIK kernel
 public:
  Pin(const std::string& path = "", const std::string& frame_id = "fr3_link8", bool urdf = true) : path(path), frame_id(frame_id), urdf(urdf) {}
  void bind(const std::shared_ptr<Sim>& s, const Pose& cfg_tcp_) { sim = s; cfg_tcp = cfg_tcp_; }
  std::optional<VectorXd> inverse(const Pose& pose, const VectorXd& q0, const Pose& tcp_offset = Pose()) override {
    if (!sim) throw std::runtime_error("Pin is not bound to a Sim yet (construct SimRobot(sim, ik, cfg) first)");
    const BatchInfo& I = sim->info;
    // the kernel applies the robot config's tcp offset; the reference drives the frame to pose * tcp_offset^-1
    const Pose goal = pose * tcp_offset.inverse() * cfg_tcp;
    const Vec3 t = goal.translation(); const Vec4 q = goal.rotation_q();
    std::vector<double> p((size_t)I.n * 7), q0s((size_t)I.n * I.nj, 0.0), out((size_t)I.n * I.ik_nq);
    std::vector<int> ok(I.n), it(I.n);
    for (int e = 0; e < I.n; e++) {
      double* r = &p[(size_t)e * 7];
      r[0] = t[0]; r[1] = t[1]; r[2] = t[2]; r[3] = q[0]; r[4] = q[1]; r[5] = q[2]; r[6] = q[3];
      for (int i = 0; i < I.nj && i < (int)q0.size(); i++) q0s[(size_t)e * I.nj + i] = q0[i];
    }
    check(rcsb_ik_inverse_host(sim->batch, p.data(), q0s.data(), out.data(), ok.data(), it.data()));
    if (!ok[0]) return std::nullopt;
    return VectorXd(out.begin(), out.begin() + I.ik_nq);
  }
  Pose forward(const VectorXd&, const Pose&) override {
    throw std::runtime_error("Pin.forward: use rcs_b200.sim.Pin.forward (host kinematics of the compiled scene)");
  }
  std::string path, frame_id;
  bool urdf;
  std::shared_ptr<Sim> sim;
  Pose cfg_tcp;
};

class SimRobot : public Robot {  // src/sim/SimRobot.{h,cpp}
 public:
  SimRobot(std::shared_ptr<Sim> sim, std::shared_ptr<Kinematics> ik, SimRobotConfig cfg, bool register_convergence_callback = true)
      : sim(sim), ik(ik), cfg(cfg) {
    (void)register_convergence_callback;  // compiled into the device model (rcs_b200.batch.DeviceModel)
    if (auto pin = std::dynamic_pointer_cast<Pin>(ik)) pin->bind(sim, cfg.tcp_offset);
    state.inverse_tcp_offset = cfg.tcp_offset.inverse();
    meta = robots_meta_config(cfg.robot_type);
  }
  RobotConfig* get_config() override { return new SimRobotConfig(cfg); }  // heap copy handed to Python (SimRobot.cpp:102-106)
  bool set_config(const SimRobotConfig& c) { cfg = c; state.inverse_tcp_offset = cfg.tcp_offset.inverse(); return true; }
  RobotState* get_state() override {
    std::vector<double> sr; std::vector<int> si;
    sim->row(sr, si);
    const int nj = sim->info.nj, o = sim->info.o_tail;
    state.previous_angles.assign(sr.begin() + o + RCSB_S_PREV, sr.begin() + o + RCSB_S_PREV + nj);
    state.target_angles.assign(sr.begin() + o + RCSB_S_TARGET, sr.begin() + o + RCSB_S_TARGET + nj);
    state.ik_success = si[RCSB_I_IK_SUCCESS]; state.collision = si[RCSB_I_COLLISION];
    state.is_moving = si[RCSB_I_MOVING]; state.is_arrived = si[RCSB_I_ARRIVED];
    return new SimRobotState(state);
  }
  Pose get_cartesian_position() override {
    std::vector<double> obs((size_t)sim->info.n * RCSB_OBS_DIM);
    check(rcsb_batch_run_host(sim->batch, RCSB_RUN_OBS, 0, 0, nullptr, nullptr, 0, nullptr, nullptr, obs.data(), nullptr));
    return Pose(Vec4{obs[3], obs[4], obs[5], obs[6]}, Vec3{obs[0], obs[1], obs[2]});
  }
  void set_joint_position(const VectorXd& q) override { sim->command(RCSB_RUN_SET_JOINTS, &q, nullptr); }
  VectorXd get_joint_position() override {
    std::vector<double> obs((size_t)sim->info.n * RCSB_OBS_DIM);
    check(rcsb_batch_run_host(sim->batch, RCSB_RUN_OBS, 0, 0, nullptr, nullptr, 0, nullptr, nullptr, obs.data(), nullptr));
    return VectorXd(obs.begin() + 7, obs.begin() + 7 + sim->info.nj);
  }
  void move_home() override { set_joint_position(meta.q_home); }
  void reset() override { check(rcsb_robot_reset(sim->batch)); }
  void close() override {}
  void set_joints_hard(const VectorXd& q) { sim->command(RCSB_RUN_SET_JOINTS_HARD, &q, nullptr); }
  void set_cartesian_position(const Pose& pose) override {
    const Vec3 t = pose.translation(); const Vec4 q = pose.rotation_q();
    std::vector<double> p((size_t)sim->info.n * 7);
    for (int e = 0; e < sim->info.n; e++) {
      double* r = &p[(size_t)e * 7];
      r[0] = t[0]; r[1] = t[1]; r[2] = t[2]; r[3] = q[0]; r[4] = q[1]; r[5] = q[2]; r[6] = q[3];
    }
    check(rcsb_robot_set_cartesian_position_host(sim->batch, p.data()));
  }
  std::optional<std::shared_ptr<Kinematics>> get_ik() override { return ik; }
  Pose get_base_pose_in_world_coordinates() override { return base_pose; }
  void set_base_pose(const Pose& p) { base_pose = p; }  // the Python layer passes the compiled scene's base frame
  std::shared_ptr<Sim> sim;
  std::shared_ptr<Kinematics> ik;
  SimRobotConfig cfg;
  SimRobotState state;
  RobotMetaConfig meta;
  Pose base_pose;
};

struct SimGripperConfig : GripperConfig {  // src/sim/SimGripper.h:15-45
  double epsilon_inner = 0.005, epsilon_outer = 0.005, seconds_between_callbacks = 0.05;
  double max_actuator_width = 255, min_actuator_width = 0, max_joint_width = 0.04, min_joint_width = 0.0;
  std::vector<std::string> ignored_collision_geoms{};
  std::vector<std::string> collision_geoms{"hand_c", "d435i_collision", "finger_0_left", "finger_0_right"};
  std::vector<std::string> collision_geoms_fingers{"finger_0_left", "finger_0_right"};
  std::string joint = "finger_joint1", actuator = "actuator8";
  void add_id(const std::string& id) {
    for (auto& s : ignored_collision_geoms) s += "_" + id;
    for (auto& s : collision_geoms) s += "_" + id;
    for (auto& s : collision_geoms_fingers) s += "_" + id;
    joint += "_" + id;
    actuator += "_" + id;
  }
};
struct SimGripperState : GripperState {  // SimGripper.h:47-52
  double last_commanded_width = 0, last_width = 0;
  bool is_moving = false, collision = false;
};
class SimGripper : public Gripper {  // src/sim/SimGripper.{h,cpp}
 public:
  SimGripper(std::shared_ptr<Sim> sim, SimGripperConfig cfg) : sim(sim), cfg(cfg) {}
  GripperConfig* get_config() override { return new SimGripperConfig(cfg); }
  bool set_config(const SimGripperConfig& c) { cfg = c; return true; }
  GripperState* get_state() override {
    std::vector<double> sr; std::vector<int> si;
    sim->row(sr, si);
    SimGripperState* s = new SimGripperState();
    s->last_commanded_width = sr[sim->info.o_tail + RCSB_S_GLCW]; s->last_width = sr[sim->info.o_tail + RCSB_S_GLW];
    s->is_moving = si[RCSB_I_G_MOVING]; s->collision = si[RCSB_I_G_COLLISION];
    return s;
  }
  void set_normalized_width(double width, double force = 0) override {
    if (width < 0 || width > 1 || force < 0) throw std::invalid_argument("width must be between 0 and 1, force must be positive");  // SimGripper.cpp:80-83
    last_commanded = width;
    sim->command(RCSB_RUN_SET_GRIPPER, nullptr, &width);
  }
  double get_normalized_width() override {
    std::vector<double> obs((size_t)sim->info.n * RCSB_OBS_DIM);
    check(rcsb_batch_run_host(sim->batch, RCSB_RUN_OBS, 0, 0, nullptr, nullptr, 0, nullptr, nullptr, obs.data(), nullptr));
    return obs[21];
  }
  bool is_grasped() override {  // SimGripper.cpp:132-141
    const double w = get_normalized_width();
    return last_commanded - cfg.epsilon_inner < w && w < last_commanded + cfg.epsilon_outer;
  }
  void grasp() override { shut(); }
  void open() override { set_normalized_width(1); }
  void shut() override { set_normalized_width(0); }
  void reset() override { check(rcsb_gripper_reset(sim->batch)); last_commanded = 0; }
  void close() override {}
  std::shared_ptr<Sim> sim;
  SimGripperConfig cfg;
  double last_commanded = 0;
};
}  // namespace sim

PYBIND11_MODULE(_core, m) {
  m.doc() = "rcs_b200._core: compiled interface of the B200 batched backend, shaped like rcs._core (src/pybind/rcs.cpp)";
  m.attr("__version__") = "0.2.0";
  auto common = m.def_submodule("common", "common module");
  common.def("IdentityTranslation", [] { auto v = IdentityTranslation(); return arr_out(v.data(), {3}); });
  common.def("IdentityRotMatrix", [] { auto v = IdentityRotMatrix(); return arr_out(v.data(), {3, 3}); });
  common.def("IdentityRotQuatVec", [] { auto v = IdentityRotQuatVec(); return arr_out(v.data(), {4}); });
  common.def("FrankaHandTCPOffset", [] { auto v = FrankaHandTCPOffset(); return arr_out(v.data(), {4, 4}); });

  py::class_<RPY>(common, "RPY")
      .def(py::init<double, double, double>(), py::arg("roll") = 0.0, py::arg("pitch") = 0.0, py::arg("yaw") = 0.0)
      .def(py::init([](const NpArr& v) { return RPY(arr_in<3>(v, "rpy")); }), py::arg("rpy"))
      .def_readwrite("roll", &RPY::roll).def_readwrite("pitch", &RPY::pitch).def_readwrite("yaw", &RPY::yaw)
      .def("rotation_matrix", [](const RPY& r) { auto v = r.rotation_matrix(); return arr_out(v.data(), {3, 3}); })
      .def("as_vector", [](const RPY& r) { auto v = r.as_vector(); return arr_out(v.data(), {3}); })
      .def("as_quaternion_vector", [](const RPY& r) { auto v = r.as_quaternion_vector(); return arr_out(v.data(), {4}); })
      .def("is_close", &RPY::is_close, py::arg("other"), py::arg("eps") = 1e-8)
      .def("__str__", &RPY::str)
      .def(py::self + py::self)
      .def(py::pickle([](const RPY& p) { return py::make_tuple(p.roll, p.pitch, p.yaw); },
                      [](py::tuple t) { return RPY(t[0].cast<double>(), t[1].cast<double>(), t[2].cast<double>()); }));

  py::class_<Pose>(common, "Pose")
      .def(py::init<>())
      .def(py::init([](const NpArr& m4) { return Pose(arr_in<16>(m4, "pose_matrix")); }), py::arg("pose_matrix"))
      .def(py::init([](const NpArr& r, const NpArr& t) { return Pose(arr_in<9>(r, "rotation"), arr_in<3>(t, "translation")); }),
           py::arg("rotation"), py::arg("translation"))
      .def(py::init([](const NpArr& q, const NpArr& t) { return Pose(arr_in<4>(q, "quaternion"), arr_in<3>(t, "translation")); }),
           py::arg("quaternion"), py::arg("translation"))
      .def(py::init([](const RPY& r, const NpArr& t) { return Pose(r, arr_in<3>(t, "translation")); }), py::arg("rpy"), py::arg("translation"))
      .def(py::init([](const NpArr& r, const NpArr& t) { return Pose::from_rpy_vector(arr_in<3>(r, "rpy_vector"), arr_in<3>(t, "translation")); }),
           py::arg("rpy_vector"), py::arg("translation"))
      .def(py::init([](const NpArr& t) { return Pose::from_translation(arr_in<3>(t, "translation")); }), py::arg("translation"))
      .def(py::init([](const NpArr& q) { return Pose(arr_in<4>(q, "quaternion")); }), py::arg("quaternion"))
      .def(py::init([](const RPY& r) { return Pose(r); }), py::arg("rpy"))
      .def(py::init([](const NpArr& r) { return Pose(arr_in<9>(r, "rotation")); }), py::arg("rotation"))
      .def(py::init([](const Pose& p) { return Pose(p); }), py::arg("pose"))
      .def("translation", [](const Pose& p) { auto v = p.translation(); return arr_out(v.data(), {3}); })
      .def("rotation_m", [](const Pose& p) { auto v = p.rotation_m(); return arr_out(v.data(), {3, 3}); })
      .def("rotation_q", [](const Pose& p) { auto v = p.rotation_q(); return arr_out(v.data(), {4}); })
      .def("pose_matrix", [](const Pose& p) { auto v = p.pose_matrix(); return arr_out(v.data(), {4, 4}); })
      .def("rotation_rpy", &Pose::rotation_rpy)
      .def("xyzrpy", [](const Pose& p) { auto v = p.xyzrpy(); return arr_out(v.data(), {6}); })
      .def("interpolate", &Pose::interpolate, py::arg("dest_pose"), py::arg("progress"))
      .def("inverse", &Pose::inverse)
      .def("total_angle", &Pose::total_angle)
      .def("limit_rotation_angle", &Pose::limit_rotation_angle, py::arg("max_angle"))
      .def("limit_translation_length", &Pose::limit_translation_length, py::arg("max_length"))
      .def("is_close", &Pose::is_close, py::arg("other"), py::arg("eps_r") = 1e-8, py::arg("eps_t") = 1e-8)
      .def("__str__", &Pose::str)
      .def(py::self * py::self)
      .def(py::pickle(
          [](const Pose& p) { auto t = p.translation(); auto q = p.rotation_q(); return py::make_tuple(q[0], q[1], q[2], q[3], t[0], t[1], t[2]); },
          [](py::tuple t) {
            return Pose::from_tq({t[4].cast<double>(), t[5].cast<double>(), t[6].cast<double>()},
                                 {t[0].cast<double>(), t[1].cast<double>(), t[2].cast<double>(), t[3].cast<double>()});
          }));

  py::class_<Kinematics, PyKinematics, std::shared_ptr<Kinematics>>(common, "Kinematics")
      .def(py::init<>())
      .def("inverse", &Kinematics::inverse, py::arg("pose"), py::arg("q0"), py::arg("tcp_offset") = Pose())
      .def("forward", &Kinematics::forward, py::arg("q0"), py::arg("tcp_offset"));
  py::class_<sim::Pin, Kinematics, std::shared_ptr<sim::Pin>>(common, "Pin")
      .def(py::init<const std::string&, const std::string&, bool>(), py::arg("path") = "", py::arg("frame_id") = "fr3_link8", py::arg("urdf") = true);

  py::enum_<RobotType>(common, "RobotType").value("FR3", FR3).value("UR5e", UR5e).value("SO101", SO101).value("XArm7", XArm7).export_values();
  py::enum_<RobotPlatform>(common, "RobotPlatform").value("HARDWARE", HARDWARE).value("SIMULATION", SIMULATION).export_values();
  py::class_<RobotMetaConfig>(common, "RobotMetaConfig")
      .def_property_readonly("q_home", [](const RobotMetaConfig& c) { return arr_out(c.q_home.data(), {(py::ssize_t)c.q_home.size()}); })
      .def_readonly("dof", &RobotMetaConfig::dof)
      .def_property_readonly("joint_limits", [](const RobotMetaConfig& c) {
        py::array_t<double> a({(py::ssize_t)2, (py::ssize_t)c.dof});
        for (int r = 0; r < 2; r++) std::memcpy(a.mutable_data(r, 0), c.joint_limits[r].data(), c.dof * sizeof(double));
        return a;
      });
  common.def("robots_meta_config", [](RobotType t) { return robots_meta_config(t); }, py::arg("robot_type"));
  py::class_<RobotConfig>(common, "RobotConfig")
      .def(py::init<>())
      .def_readwrite("robot_type", &RobotConfig::robot_type).def_readwrite("kinematic_model_path", &RobotConfig::kinematic_model_path)
      .def_readwrite("attachment_site", &RobotConfig::attachment_site).def_readwrite("tcp_offset", &RobotConfig::tcp_offset)
      .def_readwrite("robot_platform", &RobotConfig::robot_platform);
  py::class_<RobotState>(common, "RobotState");
  py::class_<GripperConfig>(common, "GripperConfig");
  py::class_<GripperState>(common, "GripperState");
  py::class_<Robot, PyRobot, std::shared_ptr<Robot>>(common, "Robot")
      .def(py::init<>())
      .def("get_config", &Robot::get_config).def("get_state", &Robot::get_state)
      .def("get_cartesian_position", &Robot::get_cartesian_position)
      .def("set_joint_position", &Robot::set_joint_position, py::arg("q"))
      .def("get_joint_position", [](Robot& r) { auto q = r.get_joint_position(); return arr_out(q.data(), {(py::ssize_t)q.size()}); })
      .def("move_home", &Robot::move_home)
      .def("reset", &Robot::reset).def("close", &Robot::close)
      .def("set_cartesian_position", &Robot::set_cartesian_position, py::arg("pose"))
      .def("get_ik", &Robot::get_ik)
      .def("get_base_pose_in_world_coordinates", &Robot::get_base_pose_in_world_coordinates)
      .def("to_pose_in_robot_coordinates", &Robot::to_pose_in_robot_coordinates, py::arg("pose_in_world_coordinates"))
      .def("to_pose_in_world_coordinates", &Robot::to_pose_in_world_coordinates, py::arg("pose_in_robot_coordinates"));
  py::class_<Gripper, PyGripper, std::shared_ptr<Gripper>>(common, "Gripper")
      .def(py::init<>())
      .def("get_config", &Gripper::get_config).def("get_state", &Gripper::get_state)
      .def("set_normalized_width", &Gripper::set_normalized_width, py::arg("width"), py::arg("force") = 0)
      .def("get_normalized_width", &Gripper::get_normalized_width)
      .def("grasp", &Gripper::grasp)
      .def("is_grasped", &Gripper::is_grasped)
      .def("open", &Gripper::open)
      .def("shut", &Gripper::shut)
      .def("close", &Gripper::close)
      .def("reset", &Gripper::reset);

  auto sm = m.def_submodule("sim", "sim module");
  py::class_<sim::SimConfig>(sm, "SimConfig")
      .def(py::init<>())
      .def_readwrite("async_control", &sim::SimConfig::async_control).def_readwrite("realtime", &sim::SimConfig::realtime)
      .def_readwrite("frequency", &sim::SimConfig::frequency).def_readwrite("max_convergence_steps", &sim::SimConfig::max_convergence_steps);
  py::class_<sim::Sim, std::shared_ptr<sim::Sim>>(sm, "Sim")
      .def(py::init<std::uintptr_t, std::uintptr_t>(), py::arg("mjmdl"), py::arg("mjdata"))
      .def("step_until_convergence", &sim::Sim::step_until_convergence, py::call_guard<py::gil_scoped_release>())
      .def("is_converged", &sim::Sim::is_converged)
      .def("step", &sim::Sim::step, py::arg("k"))
      .def("set_config", &sim::Sim::set_config, py::arg("cfg"))
      .def("get_config", &sim::Sim::get_config)
      .def("reset", &sim::Sim::reset);
  py::class_<sim::SimRobotConfig, RobotConfig>(sm, "SimRobotConfig")
      .def(py::init<>())
      .def_readwrite("joint_rotational_tolerance", &sim::SimRobotConfig::joint_rotational_tolerance)
      .def_readwrite("seconds_between_callbacks", &sim::SimRobotConfig::seconds_between_callbacks)
      .def_readwrite("mjcf_scene_path", &sim::SimRobotConfig::mjcf_scene_path)
      .def_readwrite("trajectory_trace", &sim::SimRobotConfig::trajectory_trace)
      .def_readwrite("arm_collision_geoms", &sim::SimRobotConfig::arm_collision_geoms)
      .def_readwrite("joints", &sim::SimRobotConfig::joints).def_readwrite("actuators", &sim::SimRobotConfig::actuators)
      .def_readwrite("base", &sim::SimRobotConfig::base)
      .def("add_id", &sim::SimRobotConfig::add_id, py::arg("id"));
  py::class_<sim::SimRobotState, RobotState>(sm, "SimRobotState")
      .def(py::init<>())
      .def_readonly("previous_angles", &sim::SimRobotState::previous_angles).def_readonly("target_angles", &sim::SimRobotState::target_angles)
      .def_readonly("inverse_tcp_offset", &sim::SimRobotState::inverse_tcp_offset).def_readonly("ik_success", &sim::SimRobotState::ik_success)
      .def_readonly("collision", &sim::SimRobotState::collision).def_readonly("is_moving", &sim::SimRobotState::is_moving)
      .def_readonly("is_arrived", &sim::SimRobotState::is_arrived);
  py::class_<sim::SimRobot, Robot, std::shared_ptr<sim::SimRobot>>(sm, "SimRobot")
      .def(py::init<std::shared_ptr<sim::Sim>, std::shared_ptr<Kinematics>, sim::SimRobotConfig, bool>(), py::arg("sim"), py::arg("ik"),
           py::arg("cfg"), py::arg("register_convergence_callback") = true)
      .def("get_config", &sim::SimRobot::get_config).def("set_config", &sim::SimRobot::set_config, py::arg("cfg"))
      .def("set_joints_hard", &sim::SimRobot::set_joints_hard, py::arg("q"))
      .def("_set_base_pose", &sim::SimRobot::set_base_pose, py::arg("pose"))
      .def("get_state", &sim::SimRobot::get_state);
  py::class_<sim::SimGripperConfig, GripperConfig>(sm, "SimGripperConfig")
      .def(py::init<>())
      .def_readwrite("epsilon_inner", &sim::SimGripperConfig::epsilon_inner).def_readwrite("epsilon_outer", &sim::SimGripperConfig::epsilon_outer)
      .def_readwrite("seconds_between_callbacks", &sim::SimGripperConfig::seconds_between_callbacks)
      .def_readwrite("ignored_collision_geoms", &sim::SimGripperConfig::ignored_collision_geoms)
      .def_readwrite("collision_geoms", &sim::SimGripperConfig::collision_geoms)
      .def_readwrite("collision_geoms_fingers", &sim::SimGripperConfig::collision_geoms_fingers)
      .def_readwrite("joint", &sim::SimGripperConfig::joint).def_readwrite("actuator", &sim::SimGripperConfig::actuator)
      .def_readwrite("max_actuator_width", &sim::SimGripperConfig::max_actuator_width)
      .def_readwrite("min_actuator_width", &sim::SimGripperConfig::min_actuator_width)
      .def_readwrite("max_joint_width", &sim::SimGripperConfig::max_joint_width).def_readwrite("min_joint_width", &sim::SimGripperConfig::min_joint_width)
      .def("add_id", &sim::SimGripperConfig::add_id, py::arg("id"));
  py::class_<sim::SimGripperState, GripperState>(sm, "SimGripperState")
      .def(py::init<>())
      .def_readonly("last_commanded_width", &sim::SimGripperState::last_commanded_width).def_readonly("last_width", &sim::SimGripperState::last_width)
      .def_readonly("is_moving", &sim::SimGripperState::is_moving).def_readonly("collision", &sim::SimGripperState::collision);
  py::class_<sim::SimGripper, Gripper, std::shared_ptr<sim::SimGripper>>(sm, "SimGripper")
      .def(py::init<std::shared_ptr<sim::Sim>, const sim::SimGripperConfig&>(), py::arg("sim"), py::arg("cfg"))
      .def("get_config", &sim::SimGripper::get_config).def("get_state", &sim::SimGripper::get_state)
      .def("set_config", &sim::SimGripper::set_config, py::arg("cfg"));
}
