// forwards to the single-header HDK stand-in (test infrastructure, see ../mock_hdk.h)
#pragma once
#include "../mock_hdk.h"
