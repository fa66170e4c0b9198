// tgx_trajectories.hpp — C++ host side of the drop-in: GPU-backed Circle / Line / Figure8 / Boomerang and the
// constant-speed polyline family Square / Rectangle / Reciprocating / Bounce / M / I / T.
//
// These classes keep the reference's interface — the constructor argument lists of Circle.hpp:30-31,
// Line.hpp:30-31 and Figure8.hpp:30-31 and the three overrides of the abstract Trajectory interface
// (Trajectory.hpp:33-46) — so TrajectoryGenerator.cpp:176-260 (make_unique<Circle|Line|Figure8>), :71
// (generateTraj), :419 (trajectoryInsideBounds) and :516 (generateStopTraj) compile against them unchanged.
// Everything numeric happens on the GPU behind the C-ABI in tgx.h; this layer only converts between the
// reference's containers (std::vector<Goal>, std::unordered_map<int,std::string>) and the engine's SoA planes.
//
// It derives from the reference's OWN, unmodified Trajectory.hpp, which must be on the include path (it is part of
// the reference package this library plugs into).  By default the classes live in namespace trajectory_generator
// under the reference's class names and REPLACE Circle.hpp / Line.hpp / Figure8.hpp (+ their .cpp files); define
// TGX_DROPIN_NAMESPACE to another name to place them beside the originals (the parity test does).
#pragma once

#include <string>
#include <unordered_map>
#include <vector>

#include "trajectory_generator_ros2/trajectories/Trajectory.hpp"

#include "tgx.h"

#ifndef TGX_DROPIN_NAMESPACE
#define TGX_DROPIN_NAMESPACE trajectory_generator
#endif

namespace TGX_DROPIN_NAMESPACE {

// Common implementation of the three overrides; one tgx_params record per object (plus continuation records for
// a Circle / Figure8 with more than 8 goal speeds).
class GpuTrajectory : public ::trajectory_generator::Trajectory {
public:
    ~GpuTrajectory() override;

    void generateTraj(std::vector<snapstack_msgs2::msg::Goal>& goals,
                      std::unordered_map<int, std::string>& index_msgs,
                      const rclcpp::Clock::SharedPtr& clock) override;

    void generateStopTraj(std::vector<snapstack_msgs2::msg::Goal>& goals,
                          std::unordered_map<int, std::string>& index_msgs,
                          int& pub_index,
                          const rclcpp::Clock::SharedPtr& clock) override;

    bool trajectoryInsideBounds(double xmin, double xmax,
                                double ymin, double ymax,
                                double zmin, double zmax) override;

    // Status bits (tgx_status_bits) of the last generateTraj / generateStopTraj call.
    uint32_t lastStatus() const { return last_status_; }

protected:
    GpuTrajectory(const tgx_params& params, const char* shape, const char* logger_name);
    // Circle / Figure8 with more than 8 goal speeds: `records` = the record and its continuation records (tgx.h:
    // TGX_VGOALS_MORE), exactly as they are handed to the engine
    GpuTrajectory(const std::vector<tgx_params>& records, const char* shape, const char* logger_name);

    // create<Shape>Goal(v, accel, theta): one setpoint from an explicit state, evaluated on the GPU.
    snapstack_msgs2::msg::Goal sampleGoal(double v, double accel, double s0, double s1) const;
    // create<Shape>Goal(x, y, v, accel, heading) of the polyline family (createBounceGoal: v = vz, z explicit).
    snapstack_msgs2::msg::Goal polylineGoal(double x, double y, double v, double accel, double heading,
                                            double z) const;

    tgx_params params_;
    std::vector<tgx_params> records_;   // params_ followed by its continuation records (one entry for most classes)
    std::string shape_;       // "Circle", "Line", "Figure8": prefix of the index_msgs texts
    rclcpp::Logger logger_;
    uint32_t last_status_ = 0;
};

class Circle : public GpuTrajectory {
public:
    Circle(double alt, double r, double cx, double cy,
           std::vector<double> v_goals, double t_traj, double accel, double dt);
    snapstack_msgs2::msg::Goal createCircleGoal(double v, double accel, double theta) const;
};

class Figure8 : public GpuTrajectory {
public:
    Figure8(double alt, double r, double cx, double cy,
            std::vector<double> v_goals, double t_traj, double accel, double dt);
    snapstack_msgs2::msg::Goal createFigure8Goal(double v, double accel, double theta) const;
};

class Line : public GpuTrajectory {
public:
    Line(double alt, Eigen::Vector3d A, Eigen::Vector3d B,
         std::vector<double> v_goals, double a1, double a3, double dt);
    snapstack_msgs2::msg::Goal createLineGoal(double last_x, double last_y,
                                              double v, double accel, double theta) const;
};

// Line out and back (Boomerang.hpp:30-31); announces itself as "Line traj: ..." like the reference does.
class Boomerang : public GpuTrajectory {
public:
    Boomerang(double alt, Eigen::Vector3d A, Eigen::Vector3d B,
              std::vector<double> v_goals, double a1, double a3, double dt);
    snapstack_msgs2::msg::Goal createLineGoal(double last_x, double last_y,
                                              double v, double accel, double theta) const;
};

// ---- constant-speed polyline family (Square.hpp:31-32, Rectangle.hpp, Reciprocating.hpp, Bounce.hpp, M.hpp, I.hpp,
//      T.hpp; constructed at TrajectoryGenerator.cpp:246, :258, :291-293, :318, :340, :362, :384) ----------------
class Square : public GpuTrajectory {
public:
    Square(double alt, double side_length, double cx, double cy, double orientation,
           std::vector<double> v_goals, double t_traj, double accel, double dt);
    snapstack_msgs2::msg::Goal createSquareGoal(double x, double y, double v, double accel, double heading) const;
};

class Rectangle : public GpuTrajectory {
public:
    Rectangle(double alt, double side_a, double side_b, double cx, double cy, double orientation,
              std::vector<double> v_goals, double t_traj, double accel, double dt);
    snapstack_msgs2::msg::Goal createRectangleGoal(double x, double y, double v, double accel, double heading) const;
};

class Reciprocating : public GpuTrajectory {
public:
    Reciprocating(double alt, Eigen::Vector3d A, Eigen::Vector3d B,
                  std::vector<double> v_goals, double a1, double a3, double t_traj, double dt);
    snapstack_msgs2::msg::Goal createReciprocatingGoal(double x, double y, double v, double accel,
                                                       double heading) const;
};

class Bounce : public GpuTrajectory {
public:
    Bounce(double cx, double cy, double Az, double Bz,
           std::vector<double> v_goals, double t_traj, double orientation, double dt);
    snapstack_msgs2::msg::Goal createBounceGoal(double x, double y, double z, double vz, double heading) const;
};

class M : public GpuTrajectory {
public:
    M(double cx, double cy, double length, double width, double alt,
      std::vector<double> v_goals, double t_traj, double orientation, double dt);
    snapstack_msgs2::msg::Goal createMGoal(double x, double y, double v, double accel, double heading) const;
};

class I : public GpuTrajectory {
public:
    I(double cx, double cy, double length, double width, double alt,
      std::vector<double> v_goals, double t_traj, double orientation, double dt);
    snapstack_msgs2::msg::Goal createIGoal(double x, double y, double v, double accel, double heading) const;
};

class T : public GpuTrajectory {
public:
    T(double cx, double cy, double length, double width, double alt,
      std::vector<double> v_goals, double t_traj, double orientation, double dt);
    snapstack_msgs2::msg::Goal createTGoal(double x, double y, double v, double accel, double heading) const;
};

// The process-wide engine the classes share (created on first use on device $TGX_DEVICE, default 0).
// Like the reference's objects it is not thread-safe.
tgx_engine* sharedEngine();

}  // namespace TGX_DROPIN_NAMESPACE
