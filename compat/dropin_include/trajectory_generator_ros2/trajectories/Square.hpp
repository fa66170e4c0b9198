// Drop-in replacement of the reference's trajectories/Square.hpp: the GPU-backed class of the same name
// (trajectory_generator::Square, trajectory_generator_ros2_b200/host/tgx_trajectories.hpp).  Put this directory in
// front of the reference's include/ on the include path and TrajectoryGenerator.cpp builds against it unchanged.
#pragma once
#include "tgx_trajectories.hpp"
