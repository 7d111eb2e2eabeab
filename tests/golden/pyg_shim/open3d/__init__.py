"""import-only stand-in (utils/scannet_utils.py imports open3d at module load)."""
