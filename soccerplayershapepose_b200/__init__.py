"""B200-native batched SMPL body-model layer (drop-in for the SMPL path of
xhuan8/SoccerPlayerShapePose).  See DESIGN.md."""
__version__ = "0.1.0"
