def get_normalized_weighting(numbers, weights, pairsWeight=None):
    """Faber-Ziman normalised pair weights (restatement of pdbparser's helper; inputs to the path)."""
    els = list(numbers.keys())
    total = float(sum(numbers.values()))
    c = {e: numbers[e] / total for e in els}
    norm = sum(c[e] * float(weights[e]) for e in els) ** 2
    out = {}
    for i, a in enumerate(els):
        for b in els[i:]:
            w = c[a] * c[b] * float(weights[a]) * float(weights[b]) / norm
            out[a + "-" + b] = 2.0 * w if a != b else w
    return out
