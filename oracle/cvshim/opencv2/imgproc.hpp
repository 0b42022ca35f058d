// TEST INFRASTRUCTURE — see core.hpp in this shim.
#pragma once
#include "core.hpp"
