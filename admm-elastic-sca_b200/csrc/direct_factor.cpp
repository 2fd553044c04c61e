// direct_factor.cpp -- host supernodal Cholesky (to be filled in)
