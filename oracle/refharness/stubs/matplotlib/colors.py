class ListedColormap(object):
    def __init__(self, colors):
        self.colors = colors
        self.N = len(colors)


class BoundaryNorm(object):
    def __init__(self, boundaries, ncolors):
        self.boundaries = boundaries
        self.N = ncolors
