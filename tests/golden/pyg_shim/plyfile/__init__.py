"""import-only stand-in (preprocessing/graph_level_generation.py imports plyfile at module load; the functions the
golden scripts call -- vertex_clustering, edges_from_faces -- do not use it)."""


class PlyData:  # pragma: no cover
    @staticmethod
    def read(*a, **k):
        raise NotImplementedError("plyfile is not installed in this image")
