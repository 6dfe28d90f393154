"""B200-native self-play hot path for turn-based TETR.IO (drop-in for the data-generation
path of mat-lee/tetris-reinforcement-learning).  See DESIGN.md / INTEGRATION.md."""

__version__ = "0.1.0"
