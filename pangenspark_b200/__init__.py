"""drlz-b200: B200-native drop-in for PanGenSpark's distributed RLZ parse."""
