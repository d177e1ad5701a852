// TEST INFRASTRUCTURE shim.
#pragma once
#include <string>
namespace ros { struct Time { double t = 0; static Time now() { return Time(); } }; }
namespace std_msgs { struct Header { unsigned seq = 0; ros::Time stamp; std::string frame_id; }; }
