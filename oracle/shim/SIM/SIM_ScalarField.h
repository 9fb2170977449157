// TEST INFRASTRUCTURE ONLY: forwards to the HDK node stand-in (see ../hdk_node_shim.h).
#pragma once
#include "../hdk_node_shim.h"
