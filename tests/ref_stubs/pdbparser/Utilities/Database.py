_Z = {"h": 1, "c": 6, "n": 7, "o": 8, "si": 14, "ti": 22, "ni": 28, "zr": 40}
def is_element_property(prop):
    return prop in ("atomicNumber",)
def get_element_property(element, prop):
    return _Z[str(element).lower()]
