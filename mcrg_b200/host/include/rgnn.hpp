// Drop-in replacement for the reference's src/rgnn.hpp: same names and signatures, implemented over libmcrg_b200.so.
// All declarations live in mcrg_dropin.hpp.
#pragma once
#include "mcrg_dropin.hpp"
