// TEST INFRASTRUCTURE shim.
#pragma once
#include <string>
#include <vector>
#include <std_msgs/Header.h>
namespace sensor_msgs {
struct Image { std_msgs::Header header; unsigned height = 0, width = 0, step = 0; std::string encoding; std::vector<unsigned char> data; };
namespace image_encodings { static const std::string TYPE_8UC1 = "8UC1"; static const std::string TYPE_8UC3 = "8UC3"; }
}
