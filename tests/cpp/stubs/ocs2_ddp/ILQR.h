#pragma once
#include "../ocs2_standins.h"  // stand-in of ocs2_ddp/include/ocs2_ddp/ILQR.h
