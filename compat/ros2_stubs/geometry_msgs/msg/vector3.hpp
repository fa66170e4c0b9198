// Minimal stand-in for geometry_msgs/msg/Vector3, written for this repo (NOT ROS code).
#pragma once
namespace geometry_msgs { namespace msg { struct Vector3 { double x = 0.0, y = 0.0, z = 0.0; }; } }
