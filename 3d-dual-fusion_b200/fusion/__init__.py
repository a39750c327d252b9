"""3D-DF fusion encoder (the reference's ``models/model_utils`` "ACTR" code) and the detector-side
wrappers around it, with the reference's module / parameter names so its checkpoints load."""
from .actr import ACTR, build as build_actr  # noqa: F401
from .ms_deform_attn import MSDeformAttn  # noqa: F401


def structurally_unused_parameters(model):
    """Names of parameters that can never receive a gradient on the 3D-DF path: ``level_embed`` feeds
    a position embedding no encoder layer reads, and the last hybrid layer's ``a_conv1d`` gates an
    image stream that is discarded (the reference needs
    ``find_unused_parameters=True`` for them). A DDP harness freezes these instead."""
    names = []
    for mod_name, mod in model.named_modules():
        if mod.__class__.__name__ == "DeformableTransformerACTR":
            prefix = mod_name + "." if mod_name else ""
            names.append(prefix + "level_embed")
            layers = mod.encoder.layers
            last = layers[len(layers) - 1]
            if hasattr(last, "fusion_layer") and not getattr(last, "gate_first", False):
                for n, _ in last.fusion_layer.a_conv1d.named_parameters():
                    names.append("%sencoder.layers.%d.fusion_layer.a_conv1d.%s" % (prefix, len(layers) - 1, n))
    return names
