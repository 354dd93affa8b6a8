/* case shim: the reference spells "imageLoadingUtility.h"; the file on disk is ImageLoadingUtility.h */
#pragma once
#include "ImageLoadingUtility.h"
