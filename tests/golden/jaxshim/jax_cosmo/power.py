def linear_matter_power(*a, **k):
    raise NotImplementedError("Eisenstein-Hu power is outside the hot path; goldens pass a (k, P) table")
