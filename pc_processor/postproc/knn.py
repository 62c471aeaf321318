from pmf_b200.postproc import KNN, get_gaussian_kernel  # noqa: F401
