#pragma once
#include <comn_pkg/msgs.h>
