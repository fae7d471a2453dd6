"""Import stub: the reference imports open3d for PLY IO only (custom_dataset.py:4); fixtures use .npy."""
