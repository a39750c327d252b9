"""Registry names of the config contract.

mmcv / mmdet3d / det3d are not part of this image, so a small Registry with the mmcv surface used
by the reference configs (``register_module()``, ``build(cfg)``, ``get``) is provided; when mmcv is
importable the same classes are additionally registered into its registries by ``install()``.
"""
import inspect


class Registry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def __contains__(self, key):
        return key in self._module_dict

    def __len__(self):
        return len(self._module_dict)

    def get(self, key):
        return self._module_dict.get(key)

    def _register(self, cls, name=None, force=False):
        names = [name or cls.__name__] if not isinstance(name, (list, tuple)) else list(name)
        for n in names:
            if not force and n in self._module_dict and self._module_dict[n] is not cls:
                raise KeyError("%s is already registered in %s" % (n, self._name))
            self._module_dict[n] = cls

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._register(module, name, force)
            return module
        if inspect.isclass(name):  # used as a bare decorator: @REG.register_module
            self._register(name)
            return name

        def deco(cls):
            self._register(cls, name, force)
            return cls

        return deco

    def build(self, cfg, default_args=None):
        return build_from_cfg(cfg, self, default_args)


def build_from_cfg(cfg, registry, default_args=None):
    if not isinstance(cfg, dict) or "type" not in cfg:
        raise KeyError("cfg must be a dict with the key 'type', got %r" % (cfg,))
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    obj_type = args.pop("type")
    if isinstance(obj_type, str):
        cls = registry.get(obj_type)
        if cls is None:
            raise KeyError("%s is not in the %s registry" % (obj_type, registry.name))
    else:
        cls = obj_type
    return cls(**args)


# mmdet3d-side registries (TransFusion/mmdet3d/models/registry.py, builder.py:61) and mmcv's
# CONV_LAYERS / NORM_LAYERS which the sparse blocks go through.
CONV_LAYERS = Registry("conv layer")
NORM_LAYERS = Registry("norm layer")
FUSION_LAYERS = Registry("fusion_layer")
MIDDLE_ENCODERS = Registry("middle_encoder")
VOXEL_ENCODERS = Registry("voxel_encoder")
# Det3D-side (CenterPoint/det3d/models/registry.py)
FUSION = Registry("fusion")
BACKBONES = Registry("backbone")


def build_fusion_layer(cfg):
    return FUSION_LAYERS.build(cfg)


def build_middle_encoder(cfg):
    return MIDDLE_ENCODERS.build(cfg)


def build_voxel_encoder(cfg):
    return VOXEL_ENCODERS.build(cfg)


def install():
    """Mirror our entries into mmcv / mmdet3d registries when those packages exist."""
    done = []
    try:
        from mmcv.cnn import CONV_LAYERS as MMCV_CONV
        for k, v in CONV_LAYERS.module_dict.items():
            MMCV_CONV.register_module(name=k, force=True, module=v)
        done.append("mmcv.CONV_LAYERS")
    except Exception:
        pass
    try:
        from mmdet3d.models import builder as _b
        for src, dst in ((FUSION_LAYERS, "FUSION_LAYERS"), (MIDDLE_ENCODERS, "MIDDLE_ENCODERS"),
                         (VOXEL_ENCODERS, "VOXEL_ENCODERS")):
            reg = getattr(_b, dst, None)
            if reg is not None:
                for k, v in src.module_dict.items():
                    reg.register_module(name=k, force=True, module=v)
                done.append("mmdet3d." + dst)
    except Exception:
        pass
    return done
