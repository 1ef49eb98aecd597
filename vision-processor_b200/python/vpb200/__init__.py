"""vpb200 -- host-side Python for the B200-native vision-processor detection path."""
