class pdbparser(object):
    """placeholder: the golden generators build the engine arrays themselves"""
    pass
