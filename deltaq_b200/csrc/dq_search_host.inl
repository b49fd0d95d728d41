// placeholder until the search kernels land
