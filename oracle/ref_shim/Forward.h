/* case shim: the reference spells "Forward.h" (Windows); the file on disk is forward.h */
#pragma once
#include "forward.h"
