"""eks_b200 -- B200-native (sm_100a) implementation of the EKS smoothing hot path.

Drop-in mirror of the reference's public entry points for this path:
``ensemble``, ``run_kalman_smoother``, ``optimize_smooth_param``, ``MarkerArray``,
``ensemble_kalman_smoother_singlecam`` / ``fit_eks_singlecam`` ...  All numerical work runs in
hand-written CUDA kernels behind the C ABI of ``libeks_b200.so`` (include/eks_b200.h); there is no CPU
fallback.
"""

from eks_b200.core import (  # noqa: F401
    PinholeProjection,
    ensemble,
    get_precision,
    optimize_smooth_param,
    run_kalman_smoother,
    set_precision,
)
from eks_b200.ibl_paw_multicam_smoother import fit_eks_multicam_ibl_paw  # noqa: F401
from eks_b200.ibl_pupil_smoother import (  # noqa: F401
    ensemble_kalman_smoother_ibl_pupil,
    fit_eks_pupil,
)
from eks_b200.marker_array import MarkerArray, input_dfs_to_markerArray  # noqa: F401
from eks_b200.multicam_smoother import (  # noqa: F401
    ensemble_kalman_smoother_multicam,
    fit_eks_mirrored_multicam,
    fit_eks_multicam,
)
from eks_b200.singlecam_smoother import (  # noqa: F401
    ensemble_kalman_smoother_singlecam,
    fit_eks_singlecam,
)

__version__ = '0.1.0'
