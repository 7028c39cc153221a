def print(fmt, *a, **k):
    pass
