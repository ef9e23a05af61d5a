// oracle/shim: forwarding header (test infrastructure only)
#pragma once
#include "tbb.h"
