/* case shim for "Window.h": the window / GLFW layer is not on the path; SwapChain.h only needs the name to exist */
#pragma once
