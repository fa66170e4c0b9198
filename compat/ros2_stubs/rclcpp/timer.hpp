// <rclcpp/timer.hpp>: everything lives in the rclcpp.hpp stand-in of this directory.
#pragma once
#include "rclcpp/rclcpp.hpp"
