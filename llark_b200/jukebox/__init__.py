"""B200-native mirror of the reference's jukebox/ front end (jukebox/main.py) -- see main.py and model.py."""
