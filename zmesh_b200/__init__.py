"""zmesh_b200 -- B200-native (sm_100a) multi-label marching cubes behind zmesh's `Mesher` API.

    from zmesh_b200 import Mesher
    mesher = Mesher((4, 4, 40))
    mesher.mesh(labels, close=False)
    for label in mesher.ids():
        mesh = mesher.get(label, normals=False, reduction_factor=0, voxel_centered=False)

The CUDA library (zmesh_b200/libzmesh_b200.so, C ABI in include/zmesh_b200.h) is required; there
is no CPU fallback.
"""
from .mesh import Mesh
from .mesher import Mesher
from .multi import MultiDeviceMesher

__all__ = ["Mesh", "Mesher", "MultiDeviceMesher"]
__version__ = "0.1.0"
