// TEST INFRASTRUCTURE ONLY: forwards to the HDK stand-in (see ../hdk_shim.h).
#pragma once
#include "../hdk_shim.h"
