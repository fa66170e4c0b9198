// Placeholder for snapstack_msgs2/msg/State: the trajectory samplers include it but use nothing from it.
#pragma once
namespace snapstack_msgs2 { namespace msg { struct State {}; } }
