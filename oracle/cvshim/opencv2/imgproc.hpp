// TEST INFRASTRUCTURE: forwards to the OpenCV stand-in (see cvshim.hpp).
#pragma once
#include "../cvshim.hpp"
