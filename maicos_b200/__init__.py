"""B200-native implementation of the MAICoS structure-factor / profile hot path."""
