"""edyn_b200 -- B200-native replacement for Edyn's per-step rigid-body hot path.

Host adapter (this package) -> C ABI (include/b2d.h, edyn_b200/libb2d.so) -> hand-written sm_100a kernels
(edyn_b200/csrc).  See DESIGN.md.  Importing this package never imports oracle/.
"""
from ._lib import B2DError, build  # noqa: F401
from .rigidbody import (DYNAMIC, KINEMATIC, STATIC, Material, RigidBodyDef, Shape, bodies_soa, box_shape,  # noqa: F401
                        capsule_shape, plane_shape, sphere_shape)
from .world import (World, attach, clear_collision_exclusion, detach, exclude_collision, make_hinge,  # noqa: F401
                    make_rigidbody, remove_collision_exclusion, step_simulation, update)
from . import scenes  # noqa: F401
