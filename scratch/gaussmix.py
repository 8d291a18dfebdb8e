import numpy as np
def gaussmix_formula(K=16):
    terms, names = [], []
    for k in range(1, K + 1):
        terms.append("a%d * exp(-(x - m%d)^2 / s%d^2)" % (k, k, k))
        names += ["a%d" % k, "m%d" % k, "s%d" % k]
    return " + ".join(terms), names
def gaussmix_truth(K=16):
    th = []
    for k in range(1, K + 1):
        th += [5.0 + ((7 * k) % 11), 100.0 * (k - 0.5) / K, 2.5]
    return np.array(th)
