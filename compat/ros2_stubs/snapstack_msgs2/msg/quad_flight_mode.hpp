// Placeholder for snapstack_msgs2/msg/QuadFlightMode: included by the samplers, unused by them.
#pragma once
namespace snapstack_msgs2 { namespace msg { struct QuadFlightMode {}; } }
