from scipy.linalg import cho_factor, cho_solve  # noqa: F401
