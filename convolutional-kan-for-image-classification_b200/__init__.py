"""B200-native KAN convolution hot path (source directory).  Import it as ``kanconv_b200`` (repo-root shim package);
this directory's name mirrors the upstream repository and is not a valid Python identifier."""
from ._exports import *  # noqa: F401,F403
