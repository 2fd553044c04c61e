// CollisionForce.hpp -- include-compatibility shim: the reference spreads namespace admm over several headers
// (A/src/system, A/src/collision); here everything lives in admm_b200_host.hpp.
#include "admm_b200_host.hpp"
