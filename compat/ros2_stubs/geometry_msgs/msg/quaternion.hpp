// Minimal stand-in for geometry_msgs/msg/Quaternion, written for this repo (NOT ROS code).
#pragma once
namespace geometry_msgs { namespace msg { struct Quaternion { double x = 0.0, y = 0.0, z = 0.0, w = 1.0; }; } }
