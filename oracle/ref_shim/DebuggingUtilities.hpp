// oracle/ref_shim/DebuggingUtilities.hpp -- TEST INFRASTRUCTURE: empty stand-in (tests/CameraPathGen.cpp includes it, uses nothing).
#pragma once
