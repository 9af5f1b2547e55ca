// MapPriorParams — I/algorithms/registration/map_prior.hpp:15-21.  The parameter struct lives in registration_params.hpp
// here; the prior itself (MapPrior::update / apply / prior_error, :30-146) runs inside libspx
// (spx_registration_set_map_prior_state, csrc/spx_registration.cu).
#pragma once

#include "sycl_points/algorithms/registration/registration_params.hpp"
