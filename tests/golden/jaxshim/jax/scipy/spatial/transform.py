from scipy.spatial.transform import Rotation  # noqa: F401
