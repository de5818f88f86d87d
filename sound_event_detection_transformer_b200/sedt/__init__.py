"""Mirror of the reference package `sedt`: same factory, same classes, same
call contracts (sedt/__init__.py:8-63), with the forward and the matcher
running in hand-written sm_100a kernels."""
from .criterion import PostProcess, SetCriterion
from .matcher import HungarianMatcher, build_matcher
from .model import SEDT, SPSEDT
from .modules import Backbone, Joiner, PositionEmbeddingSine, Transformer


def build_position_encoding(args):
    if args.position_embedding in ("v2", "sine"):
        return PositionEmbeddingSine(args.hidden_dim, normalize=True)
    if args.position_embedding in ("v3", "learned"):
        raise NotImplementedError("learned position embedding is not used by any documented recipe "
                                  "(SURVEY.md section 2, #3) and is not on the B200 path")
    raise ValueError(f"not supported {args.position_embedding}")


def build_backbone(args):
    """sedt/backbone.py:135-141."""
    backbone = Backbone(args.backbone, args.lr_backbone > 0, args.dilation)
    return Joiner(backbone, build_position_encoding(args))


def build_transformer(args):
    """sedt/transformer.py:409-420."""
    return Transformer(d_model=args.hidden_dim, dropout=args.dropout, nhead=args.nheads,
                       dim_feedforward=args.dim_feedforward, num_encoder_layers=args.enc_layers,
                       num_decoder_layers=args.dec_layers, normalize_before=args.pre_norm, self_sup=args.self_sup)


def build_model(args):
    """Same signature and return triple as the reference's sedt.build_model(args).
    Optional extra attributes on `args`: precision ("bf16" | "fp32"), use_tensor_cores (bool)."""
    num_classes = 1 if args.self_sup else args.num_classes
    precision = getattr(args, "precision", "bf16")
    use_tc = bool(getattr(args, "use_tensor_cores", True))
    backbone = build_backbone(args)
    transformer = build_transformer(args)
    if args.self_sup:
        model = SPSEDT(backbone, transformer, num_classes=num_classes, num_queries=args.num_queries,
                       aux_loss=args.aux_loss, feature_recon=args.feature_recon, query_shuffle=args.query_shuffle,
                       num_patches=args.num_patches, precision=precision, use_tensor_cores=use_tc)
    else:
        model = SEDT(backbone, transformer, num_classes=num_classes, num_queries=args.num_queries,
                     aux_loss=args.aux_loss, dec_at=args.dec_at, pooling=args.pooling, precision=precision,
                     use_tensor_cores=use_tc)
    matcher = build_matcher(args)
    weight_dict = {"loss_ce": args.ce_loss_coef, "loss_bbox": args.bbox_loss_coef, "loss_giou": args.giou_loss_coef}
    losses = ["labels", "boxes", "cardinality"]
    if not args.self_sup:
        if args.dec_at:
            weight_dict["loss_weak"] = args.weak_loss_coef
            losses += ["weak"]
        if args.pooling:
            weight_dict["loss_weak_p"] = args.weak_loss_p_coef
    elif args.feature_recon:
        losses += ["feature"]
        weight_dict["loss_feature"] = 1
    if args.aux_loss:
        aux = {}
        for i in range(args.dec_layers - 1):
            aux.update({f"{k}_{i}": v for k, v in weight_dict.items()})
        weight_dict.update(aux)
    criterion = SetCriterion(num_classes, matcher=matcher, weight_dict=weight_dict, eos_coef=args.eos_coef, losses=losses)
    import torch
    if torch.cuda.is_available():                 # to_cuda_if_available (utilities/utils.py:85-110)
        criterion = criterion.cuda()
    return model, criterion, {"bbox": PostProcess()}


__all__ = ["build_model", "build_matcher", "build_backbone", "build_transformer", "SEDT", "SPSEDT", "SetCriterion",
           "PostProcess", "HungarianMatcher"]
