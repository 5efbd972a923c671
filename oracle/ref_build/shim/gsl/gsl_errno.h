// empty stand-in (GSL not installed)
